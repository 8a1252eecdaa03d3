"""Training step of the hot path: forward with saved activations, backward, TF-flavoured Adam -- the B200 replacement of
``sess.run([train_op, loss, state, ...])`` (reference run_hier_xing.py:301-302; train_op =
``tf.train.AdamOptimizer(lr).minimize(loss)``, model.py:134-141; lr schedule run_hier_xing.py:265-269).

All arithmetic is in libhtcn.so (include/htcn.h, "Training step"); torch provides device memory, streams and the NCCL
all-reduce of the gradients for data-parallel training.  fp32 like the reference.

Every parameter, gradient and Adam moment lives in ONE flat fp32 buffer each, so the optimiser is a single launch and the
data-parallel exchange a single all-reduce.  The output table is kept as W_out^T [N,128] (what the scoring kernels read);
``state_dict()`` transposes it back to the TF variable layout.  Like TensorFlow's gradients through the one-hot matmuls,
the embedding / output-table gradients are dense, and Adam's moment decay touches every row every step.
"""
from __future__ import annotations

import math
import os

import numpy as np

from . import _cabi as cabi
from .model_hier import D, HierTCN, _torch


class HierTCNTrainer:
    """``Trainer(model).train_step(x_list, y_list, mask_list, state)`` == one ``sess.run([train_op, loss, state])``.

    model: a built ``HierTCN``; its parameter tensors are re-homed into the flat buffer (the model keeps working for
    evaluation and sees every update).  Parameters, gradients, Adam and the K1/K2/K3 forward+backward are fp32 in both
    tiers; with ``precision='bf16'`` the catalog products (the loss sweep and its backward, > 95% of the FLOPs) run on the
    tensor cores with bf16 operands derived from the fp32 masters after every update.  ``dist``: a torch.distributed module with an initialised process
    group for data-parallel training over users (gradients and the user count are all-reduced), or None."""

    def __init__(self, model: HierTCN, learning_rate=None, beta1=0.9, beta2=0.999, eps=1e-8, dist=None, world=1):
        torch = _torch()
        if not model.built:
            model.build()
        self.bf16 = model.precision == "bf16"
        if model.n_out != model.N:
            raise NotImplementedError("training with a catalog-sharded output table")
        if getattr(model, "wide", False):
            raise NotImplementedError("training with tcn_channel above 128 (the two-plane fp32 kernels are forward only)")
        if model.l2_normalize:
            raise NotImplementedError("training through the l2-normalised head (model_tcn.py:42-43): evaluation only")
        self.m, self.dist, self.world = model, dist, int(world)
        self.lr = float(model.args.learning_rate if learning_rate is None else learning_rate)
        self.beta1, self.beta2, self.eps = float(beta1), float(beta2), float(eps)
        self.t = 0
        m = model
        N, G, L, K = m.N, m.G, m.n_levels, m.K
        spec = [("E", (N, D)), ("wt", (N, D)), ("b_out", (N,)), ("b_emb", (D,)), ("w_in_x", (D, D)),
                ("w_in_state", (G * D, D))]
        for l in range(L):
            spec += [(f"conv_w{l}", (K, D, D)), (f"conv_b{l}", (D,))]
            if m.ds_w[l] is not None:           # 1x1 down-sample residual of a level that changes the width
                spec += [(f"ds_w{l}", (D, D)), (f"ds_b{l}", (D,))]
        for g in range(G):
            spec += [(f"gate_w{g}", (2 * D, 2 * D)), (f"gate_b{g}", (2 * D,)), (f"cand_w{g}", (2 * D, D)), (f"cand_b{g}", (D,))]
        self.spec, self.offsets, off = spec, {}, 0
        for name, shape in spec:
            self.offsets[name] = off
            off += -(-int(np.prod(shape)) // 4) * 4                 # 16-byte aligned segments
        self.n_flat = off
        f32 = torch.float32
        self.params = torch.zeros(off, dtype=f32, device=m.device)
        self.peer = None
        if dist is not None and self.world > 1 and not os.environ.get("HTCN_TRAIN_NCCL"):
            # data parallel over NVLink peer memory: the gradient buffer lives in a symmetric allocation every rank maps, and
            # all-reduce + Adam are one kernel (htcn_peer_allreduce_adam); a setup failure raises on every rank alike -> NCCL
            from .peer import PeerBuffer, _align
            try:
                o_sc = 512
                o_g = o_sc + 256
                o_r = o_g + _align(off * 4)
                self.peer = PeerBuffer(dist, dist.get_rank(), self.world, o_r + _align(off * 4), m.device)
                self._peer_off = (o_sc, o_g, o_r)
            except RuntimeError:
                self.peer = None
        if self.peer is not None:
            b = self.peer
            self.grads = b.view(self._peer_off[1], (off,), f32)
            self._sym_scalars = b.view(self._peer_off[0], (8,), f32)
            self._scalars_out = torch.zeros(8, dtype=f32, device=m.device)
            W, r = self.world, b.rank
            mk = lambda o: cabi.ptr_array([b.base[p] + o for p in range(W)])  # noqa: E731
            self._peer_args = dict(grads=mk(self._peer_off[1]), reduced=mk(self._peer_off[2]), scalars=mk(self._peer_off[0]),
                                   f_in=cabi.ptr_array([b.base[p] + (0 * 8 + r) * 4 for p in range(W)]),
                                   f_mid=cabi.ptr_array([b.base[p] + (1 * 8 + r) * 4 for p in range(W)]))
        else:
            self.grads = torch.zeros(off, dtype=f32, device=m.device)
        self.adam_m = torch.zeros(off, dtype=f32, device=m.device)
        self.adam_v = torch.zeros(off, dtype=f32, device=m.device)
        wt_master = m.wt if not self.bf16 else torch.from_numpy(
            np.ascontiguousarray(np.asarray(m._w_out_host, dtype=np.float32).T)).to(m.device)
        if m.emb_pitch != D:            # training keeps the table at the full 128-float pitch (dense dE, one flat buffer)
            E_full = torch.zeros((N, D), dtype=f32, device=m.device)
            E_full[:, :m.emb_pitch] = m.E
            m.E, m.emb_pitch = E_full, D
        cur = {"E": m.E, "wt": wt_master, "b_out": m.b_out, "b_emb": m.b_emb, "w_in_x": m.w_in_x, "w_in_state": m.w_in_state}
        for l in range(L):
            cur[f"conv_w{l}"], cur[f"conv_b{l}"] = m.conv_w[l], m.conv_b[l]
            if m.ds_w[l] is not None:
                cur[f"ds_w{l}"], cur[f"ds_b{l}"] = m.ds_w[l], m.ds_b[l]
        for g in range(G):
            cur[f"gate_w{g}"], cur[f"gate_b{g}"], cur[f"cand_w{g}"], cur[f"cand_b{g}"] = m.gru[g]
        self.p, self.g = {}, {}
        for name, shape in spec:
            self.p[name] = self._view(self.params, name, shape)
            self.g[name] = self._view(self.grads, name, shape)
            self.p[name].copy_(cur[name].reshape(shape))
        # re-home the model's tensors (views of the flat buffer: updates are visible to forward / evaluation)
        m.E, m.b_out, m.b_emb = self.p["E"], self.p["b_out"], self.p["b_emb"]
        m.wt_f32 = self.p["wt"]
        if not self.bf16:
            m.wt = self.p["wt"]
        else:       # m.wt stays the derived bf16 scoring layout [N,144]; plus the TF-layout copy the backward streams
            self.n_pad = -(-N // 8) * 8
            self.w_out_bf16 = torch.zeros((D, self.n_pad), dtype=torch.bfloat16, device=m.device)
            self._refresh_bf16_tables()
        m.w_in_x, m.w_in_state = self.p["w_in_x"], self.p["w_in_state"]
        m.conv_w = [self.p[f"conv_w{l}"] for l in range(L)]
        m.conv_b = [self.p[f"conv_b{l}"] for l in range(L)]
        m.ds_w = [self.p.get(f"ds_w{l}") for l in range(L)]
        m.ds_b = [self.p.get(f"ds_b{l}") for l in range(L)]
        m.gru = [tuple(self.p[f"{n}{g}"] for n in ("gate_w", "gate_b", "cand_w", "cand_b")) for g in range(G)]
        m.refresh_pointer_tables()
        self._d_conv_w = cabi.ptr_array([self.g[f"conv_w{l}"].data_ptr() for l in range(L)])
        self._d_conv_b = cabi.ptr_array([self.g[f"conv_b{l}"].data_ptr() for l in range(L)])
        null = (None, None)
        self._d_ds_w = cabi.ptr_array([self.g[f"ds_w{l}"].data_ptr() if f"ds_w{l}" in self.g else 0 for l in range(L)]) if m.has_ds else null
        self._d_ds_b = cabi.ptr_array([self.g[f"ds_b{l}"].data_ptr() if f"ds_b{l}" in self.g else 0 for l in range(L)]) if m.has_ds else null
        self._d_gru = [cabi.ptr_array([self.g[f"{n}{g}"].data_ptr() for g in range(G)])
                       for n in ("gate_w", "gate_b", "cand_w", "cand_b")]
        torch.cuda.synchronize(m.device)

    def _refresh_bf16_tables(self):
        m = self.m
        st = m.stream_ptr()
        cabi.call("htcn_refresh_wout", self.p["wt"].data_ptr(), self.p["b_out"].data_ptr(), m.N, m.wt.data_ptr(),
                  cabi.HTCN_BF16, st)
        cabi.call("htcn_cast_transpose_bf16", self.p["wt"].data_ptr(), cabi.HTCN_F32, m.N, None, self.w_out_bf16.data_ptr(),
                  self.n_pad, st)

    def _view(self, flat, name, shape):
        o = self.offsets[name]
        return flat[o:o + int(np.prod(shape))].view(*shape)

    # ------------------------------------------------------------------ forward + backward
    def dropout_scales(self, S, masks=None):
        """[S, n_levels, 128] device array of dropout scales for one training step, or None when args.dropout == 0.
        The reference's TemporalBlock applies tf.layers.Dropout(rate, noise_shape=[1,1,C]) to relu(conv) before the residual
        add (customized_tcn_cell.py:100,119): one Bernoulli(1-rate) channel mask per dropout op -- the unrolled graph has
        one op per (session slot, level) --, shared over batch and time, scaled by 1/(1-rate).  ``masks``: explicit
        [S, n_levels, C] 0/1 keep-masks (tests); otherwise drawn from a numpy generator seeded by (seed, step)."""
        rate = float(self.m.args.dropout)
        if rate <= 0.0 and masks is None:
            return None
        torch = _torch()
        L = self.m.n_levels
        if masks is None:
            rng = np.random.default_rng([int(self.m.seed), self.t])
            masks = rng.random((S, L, D)) >= rate
        keep = 1.0 - rate
        sc = np.zeros((S, L, D), np.float32)
        mk = np.asarray(masks, np.float32)
        sc[:, :, :mk.shape[2]] = mk / np.float32(keep)
        return torch.from_numpy(sc).to(self.m.device)

    def forward_backward(self, x_list=None, y_list=None, mask_list=None, state=None, staged=None, metrics=False,
                         mask_warmstart=None, x_gap=None, dropout_masks=None, neg_ids=None, loss_kind=None):
        """Accumulates the gradients of sum_b(sum_t loss/(n_b+1e-6)) into ``self.grads`` (the 1/user_count is applied by
        the optimiser).  Returns dict(scalars [8] device: loss, ..., user_count, n_valid; state [B,G*H] device).
        ``neg_ids [Q,k]`` (k <= 32; host or device): train on the sampled ranking loss of reference loss.py:22-71
        (``loss_kind`` or args.loss: nce / hinge_sigmoid / hinge_logsigmoid / hinge_linear / bpr) against the rows of the
        output table instead of the full-softmax cross-entropy -- no catalog sweep, gradients reach only the gathered rows."""
        torch = _torch()
        m = self.m
        d = staged if staged is not None else m.stage(x_list, y_list, mask_list, state, None, mask_warmstart, x_gap)
        y_loss = d.get("y_loss", d["y_id"])     # mask_y of model.py:62,102-103: targets with the warm-start mask applied
        m.generation += 1               # the loss workspaces are shared with HierTCN.score: older handles are stale
        B, T, S, Q = d["B"], d["T"], d["S"], d["Q"]
        G, L, K, N = m.G, m.n_levels, m.K, m.N
        R = B * T
        st = m.stream_ptr()
        f32 = torch.float32
        P = lambda t: t.data_ptr() if t is not None else None  # noqa: E731
        slot_p, slot_keep = cabi.int_array(d["slot_off"])
        buf = m._buf
        # ---- forward, keeping what the backward needs
        # The conv stack runs on the fp32 level kernels by default.  bf16 tier with ``self.k2_tcgen05 = True``: the fused
        # tcgen05 kernel (activations saved in bf16; -5 ms of 49.5 at config-5 size).  Its activations carry bf16 noise, which
        # flips the ReLU gates of ~0.2% of the near-zero pre-activations; against an exact-arithmetic gradient that is a
        # relative error of sqrt(0.002) ~ 4% in every tensor below the conv stack (3e-3 with the fp32 conv stack), unbiased.
        fused = self.bf16 and getattr(self, "k2_tcgen05", False) and (K - 1) * (1 << max(L - 1, 0)) <= 32
        sdt, sdt_c = (torch.bfloat16, cabi.HTCN_BF16) if fused else (f32, cabi.HTCN_F32)
        xe = buf("tr_xe_bf16" if fused else "tr_xe", (R, D), sdt)
        yp = buf("tr_yp", (S, B, D), f32)
        cabi.call("htcn_gather_meanpool", m.E.data_ptr(), D, m.b_emb.data_ptr(), N, d["x_id"].data_ptr(), d["y_id"].data_ptr(),
                  slot_p, B, T, S, xe.data_ptr(), sdt_c, yp.data_ptr(), st)
        state_pre = buf("tr_state_pre", (S, B, G * D), f32)
        sbias = buf("tr_sbias", (S, B, D), f32)
        gates = buf("tr_gates", (S, G, 3, B, D), f32)
        state_out = torch.empty((B, G * D), dtype=f32, device=m.device)
        if self.bf16 and G == 2 and getattr(self, "k3_tcgen05", bool(os.environ.get("HTCN_TRAIN_K3_BF16"))):
            # opt-in (``self.k3_tcgen05 = True`` / HTCN_TRAIN_K3_BF16=1): the tensor-core GRU of the inference path with the gate
            # activations written out -- 0.08 ms against 1.0 ms for the fp32 kernel at 4096 users x 10 sessions (cfg5 step 24.7 ->
            # 24.2 ms), but bf16 recurrent operands move every gradient below the GRU by 2-4 % (norm-relative, cosine >= 0.999)
            # where the fp32 kernel keeps them within 0.4 % of the fp64 oracle: not the default
            k3_scratch = buf("k3_scratch_bf16", (cabi.gru_scratch_bytes(B) // 4,), f32)
            cabi.call("htcn_gru_sessions_train_bf16", yp.data_ptr(), d["mask"].data_ptr(), d["state"].data_ptr(), m._gru_pp[0][0],
                      m._gru_pp[1][0], m._gru_pp[2][0], m._gru_pp[3][0], G, m.w_in_state.data_ptr(), B, S,
                      k3_scratch.data_ptr(), state_pre.data_ptr(), sbias.data_ptr(), state_out.data_ptr(), gates.data_ptr(), st)
        else:
            cabi.call("htcn_gru_sessions_train", yp.data_ptr(), d["mask"].data_ptr(), d["state"].data_ptr(), m._gru_pp[0][0],
                      m._gru_pp[1][0], m._gru_pp[2][0], m._gru_pp[3][0], G, m.w_in_state.data_ptr(), B, S,
                      state_pre.data_ptr(), sbias.data_ptr(), state_out.data_ptr(), gates.data_ptr(), st)
        h_save = buf("tr_h_save_bf16" if fused else "tr_h_save", (L + 1, R, D), sdt)
        a_save = buf("tr_a_save_bf16" if fused else "tr_a_save", (max(L, 1), R, D), sdt)
        hout = buf("tr_hout_bf16" if fused else "tr_hout", (max(Q, 1), D), sdt)
        drop = self.dropout_scales(S, dropout_masks)             # None unless args.dropout > 0 (training only)
        if fused:
            k2f_scratch = buf("k2_scratch_bf16", (cabi.tcn_scratch_floats(L, K),), f32)
            cabi.call("htcn_tcn_forward_train_bf16", xe.data_ptr(), m.w_in_x.data_ptr(), sbias.data_ptr(), m._conv_w_pp[0],
                      m._conv_b_pp[0], m._ds_w_pp[0], m._ds_b_pp[0], L, K, slot_p, B, T, S, d["row_of"].data_ptr(), P(drop),
                      h_save.data_ptr(), a_save.data_ptr(), hout.data_ptr(), k2f_scratch.data_ptr(), st)
            cabi.note_launches(2)
        else:
            # bf16 tier: the levels run on the tensor cores with fp32-grade split products (``self.k2_split_tc = False``: FFMA)
            tcf = None
            if self.bf16 and getattr(self, "k2_split_tc", True):
                tcf = buf("tr_k2_tcf_scratch", (cabi.K2TC_SCRATCH_BYTES,), torch.uint8)
            cabi.call("htcn_tcn_forward_train", xe.data_ptr(), m.w_in_x.data_ptr(), sbias.data_ptr(), m._conv_w_pp[0],
                      m._conv_b_pp[0], m._ds_w_pp[0], m._ds_b_pp[0], L, K, slot_p, B, T, S, d["row_of"].data_ptr(), P(drop),
                      h_save.data_ptr(), a_save.data_ptr(), hout.data_ptr(), P(tcf), st)
            cabi.note_launches(L + 2 + (L + 1 if tcf is not None else 0))
        scalars = torch.zeros(8, dtype=f32, device=m.device)
        if Q == 0:
            return dict(scalars=scalars, state=state_out)
        if neg_ids is not None:
            # ---- sampled ranking loss (loss.py:22-71): 1 + k gathered rows of W_out^T per scored position
            a = m.args
            kind = loss_kind or (a.loss if a.loss in cabi.LOSS_KINDS else "hinge_logsigmoid")
            neg = neg_ids if hasattr(neg_ids, "data_ptr") else torch.from_numpy(np.ascontiguousarray(neg_ids, np.int32)).to(m.device)
            k = int(neg.shape[1])
            assert int(neg.shape[0]) == Q, "neg_ids must have one row per scored position"
            h_prec = cabi.HTCN_BF16 if fused else cabi.HTCN_F32
            loss_row = buf("loss_row", (Q,), f32)
            cabi.call("htcn_sampled_rank_loss", hout.data_ptr(), h_prec, Q, self.p["wt"].data_ptr(), d["y_rows"].data_ptr(),
                      neg.data_ptr(), k, cabi.LOSS_KINDS[kind], float(a.hinge_delta), float(a.nce_weight),
                      int(a.num_neg_sample), loss_row.data_ptr(), st)
            cabi.call("htcn_loss_metrics_reduce", loss_row.data_ptr(), None, d["row_of"].data_ptr(), y_loss.data_ptr(),
                      B, T, N, None, None, None, buf("user_part", (B, 8), f32).data_ptr(), scalars.data_ptr(), st)
            g_row = buf("tr_g_row", (Q,), f32)
            cabi.call("htcn_loss_row_weights", y_loss.data_ptr(), d["row_of"].data_ptr(), B, T, g_row.data_ptr(), st)
            d_hout = buf("tr_d_hout", (Q, D), f32)
            cabi.call("htcn_sampled_rank_loss_backward", hout.data_ptr(), h_prec, Q, self.p["wt"].data_ptr(),
                      d["y_rows"].data_ptr(), neg.data_ptr(), k, cabi.LOSS_KINDS[kind], float(a.hinge_delta),
                      float(a.nce_weight), int(a.num_neg_sample), g_row.data_ptr(), d_hout.data_ptr(),
                      self.g["wt"].data_ptr(), st)
            return self._backward_below_head(d, d_hout, xe, sdt_c, yp, state_pre, gates, h_save, a_save, drop, scalars,
                                             state_out, slot_p, slot_keep)
        # ---- loss (one streaming sweep: log-sum-exp per row, optionally the rank metrics)
        ns = m.n_split_for(Q, N)
        zy = buf("zy", (Q,), f32)
        pm, ps = buf("pm", (ns, Q), f32), buf("ps", (ns, Q), f32)
        pc = buf("pc", (ns, Q), torch.int32) if metrics else None
        hq, prec = hout, cabi.HTCN_F32
        if self.bf16:     # bf16 user embeddings: rows for the sweeps, the transpose for dW^T = P^T Hout
            q_pad = -(-Q // 8) * 8
            hq_t = buf("tr_hout_t_bf16", (D, q_pad), torch.bfloat16)
            if fused:
                cabi.call("htcn_cast_transpose_bf16", hout.data_ptr(), cabi.HTCN_BF16, Q, None, hq_t.data_ptr(), q_pad, st)
            else:
                hq = buf("tr_hout_bf16", (Q, D), torch.bfloat16)
                cabi.call("htcn_cast_transpose_bf16", hout.data_ptr(), cabi.HTCN_F32, Q, hq.data_ptr(), hq_t.data_ptr(), q_pad, st)
            prec = cabi.HTCN_BF16
        cabi.call("htcn_target_logit", hq.data_ptr(), prec, Q, m.wt.data_ptr(), m.b_out.data_ptr(), N, 0,
                  d["y_rows"].data_ptr(), zy.data_ptr(), st)
        loss_row = buf("loss_row", (Q,), f32)
        g_row = buf("tr_g_row", (Q,), f32)
        cabi.call("htcn_loss_row_weights", y_loss.data_ptr(), d["row_of"].data_ptr(), B, T, g_row.data_ptr(), st)
        d_hout = buf("tr_d_hout", (Q, D), f32)
        if self.bf16 and not metrics and getattr(self, "k4_fused_fwd_bwd", True):
            # loss + dHout in ONE catalog sweep (target-referenced sums need no running maximum), dW^T / db in a second
            cabi.call("htcn_score_ce_fwd_bwd_bf16", hq.data_ptr(), hq_t.data_ptr(), q_pad, Q, m.wt.data_ptr(),
                      self.w_out_bf16.data_ptr(), self.n_pad, m.b_out.data_ptr(), N, 0, d["y_rows"].data_ptr(), zy.data_ptr(),
                      g_row.data_ptr(), buf("tr_k4_ws", (cabi.ce_bwd_bf16_ws_floats(Q, N),), f32).data_ptr(),
                      loss_row.data_ptr(), d_hout.data_ptr(), self.g["wt"].data_ptr(), self.g["b_out"].data_ptr(), st)
            cabi.call("htcn_loss_metrics_reduce", loss_row.data_ptr(), None, d["row_of"].data_ptr(), y_loss.data_ptr(),
                      B, T, N, None, None, None, buf("user_part", (B, 8), f32).data_ptr(), scalars.data_ptr(), st)
            return self._backward_below_head(d, d_hout, xe, sdt_c, yp, state_pre, gates, h_save, a_save, drop, scalars,
                                             state_out, slot_p, slot_keep)
        cabi.call("htcn_score_ce_rank_topk", hq.data_ptr(), prec, Q, m.wt.data_ptr(), m.b_out.data_ptr(), N, 0,
                  d["y_rows"].data_ptr(), zy.data_ptr(), 1, cabi.SCORE_CE | (cabi.SCORE_RANK if metrics else 0), 0, ns,
                  pm.data_ptr(), ps.data_ptr(), P(pc), None, None, st)
        cabi.note_launches(-1)
        rank_row = buf("rank_row", (Q,), f32) if metrics else None
        cabi.call("htcn_score_finish", pm.data_ptr(), ps.data_ptr(), P(pc), ns, Q, d["y_rows"].data_ptr(), zy.data_ptr(),
                  loss_row.data_ptr(), P(rank_row), st)
        if self.bf16:
            cabi.call("htcn_score_ce_repair", hq.data_ptr(), prec, Q, m.wt.data_ptr(), N, zy.data_ptr(), loss_row.data_ptr(),
                      None, st)
        cabi.call("htcn_loss_metrics_reduce", loss_row.data_ptr(), P(rank_row), d["row_of"].data_ptr(), y_loss.data_ptr(),
                  B, T, N, None, None, None, buf("user_part", (B, 8), f32).data_ptr(), scalars.data_ptr(), st)
        # ---- backward
        if self.bf16:
            cabi.call("htcn_score_ce_backward_bf16", hq.data_ptr(), hq_t.data_ptr(), q_pad, Q, m.wt.data_ptr(),
                      self.w_out_bf16.data_ptr(), self.n_pad, m.b_out.data_ptr(), N, 0, d["y_rows"].data_ptr(),
                      loss_row.data_ptr(), zy.data_ptr(), g_row.data_ptr(),
                      buf("tr_k4_ws", (cabi.ce_bwd_bf16_ws_floats(Q, N),), f32).data_ptr(), d_hout.data_ptr(),
                      self.g["wt"].data_ptr(), self.g["b_out"].data_ptr(), st)
        else:
            cabi.call("htcn_score_ce_backward", hout.data_ptr(), cabi.HTCN_F32, Q, m.wt.data_ptr(), m.b_out.data_ptr(), N, 0,
                      d["y_rows"].data_ptr(), loss_row.data_ptr(), zy.data_ptr(), g_row.data_ptr(), d_hout.data_ptr(),
                      self.g["wt"].data_ptr(), self.g["b_out"].data_ptr(), st)
        return self._backward_below_head(d, d_hout, xe, sdt_c, yp, state_pre, gates, h_save, a_save, drop, scalars, state_out,
                                         slot_p, slot_keep)

    def _backward_below_head(self, d, d_hout, xe, sdt_c, yp, state_pre, gates, h_save, a_save, drop, scalars, state_out,
                             slot_p, slot_keep):
        """conv stack, GRU (BPTT over the S steps) and embedding backward, given dL/dHout"""
        torch = _torch()
        m = self.m
        B, T, S = d["B"], d["T"], d["S"]
        G, L, K, N = m.G, m.n_levels, m.K, m.N
        R = B * T
        st = m.stream_ptr()
        f32 = torch.float32
        P = lambda t: t.data_ptr() if t is not None else None  # noqa: E731
        buf = m._buf
        d_sbias = buf("tr_d_sbias", (S, B, D), f32)
        d_xe = buf("tr_d_xe", (R, D), f32)
        k2_scratch = buf("tr_k2_scratch", (3 if m.has_ds else 2, R, D), f32)
        # bf16 tier: weight gradients of the conv stack on the tensor cores (set self.wgrad_tcgen05 = False for the fp32
        # split-K products); the relu gates stay those of the fp32 forward, only the products round their operands to bf16
        tc = None
        if self.bf16 and getattr(self, "wgrad_tcgen05", True):
            tc = buf("tr_k2_tc_scratch", (int(cabi.load().htcn_tcn_backward_tc_scratch_bytes(B, T, S, L, K)),), torch.uint8)
        cabi.call("htcn_tcn_backward", d_hout.data_ptr(), d["row_of"].data_ptr(), xe.data_ptr(), sdt_c, m.w_in_x.data_ptr(),
                  m._conv_w_pp[0], m._ds_w_pp[0], L, K, slot_p, B, T, S, P(drop), h_save.data_ptr(), a_save.data_ptr(),
                  k2_scratch.data_ptr(), P(tc), self._d_conv_w[0], self._d_conv_b[0], self._d_ds_w[0], self._d_ds_b[0],
                  self.g["w_in_x"].data_ptr(), d_sbias.data_ptr(), d_xe.data_ptr(), st)
        cabi.note_launches(1 + L * (K + 3) + 3)
        d_yp = buf("tr_d_yp", (S, B, D), f32)
        n_scr = int(cabi.load().htcn_gru_backward_scratch_floats(B, S, G))
        k3_scratch = buf("tr_k3_scratch", (n_scr,), f32)
        cabi.call("htcn_gru_backward", yp.data_ptr(), d["mask"].data_ptr(), state_pre.data_ptr(), gates.data_ptr(),
                  m._gru_pp[0][0], m._gru_pp[2][0], G, m.w_in_state.data_ptr(), B, S, d_sbias.data_ptr(),
                  k3_scratch.data_ptr(), self._d_gru[0][0], self._d_gru[1][0], self._d_gru[2][0], self._d_gru[3][0],
                  self.g["w_in_state"].data_ptr(), d_yp.data_ptr(), st)
        cabi.note_launches(2 + G + 9 * G)          # setup, G sbias products, ONE recurrence kernel, 9 deferred products per layer
        cabi.call("htcn_gather_backward", d_xe.data_ptr(), d_yp.data_ptr(), d["x_id"].data_ptr(), d["y_id"].data_ptr(),
                  slot_p, B, T, S, N, self.g["E"].data_ptr(), self.g["b_emb"].data_ptr(), st)
        del slot_keep
        return dict(scalars=scalars, state=state_out)

    # ------------------------------------------------------------------ optimiser
    def lr_t(self, lr=None):
        lr = self.lr if lr is None else float(lr)
        return lr * math.sqrt(1.0 - self.beta2 ** self.t) / (1.0 - self.beta1 ** self.t)

    def apply_gradients(self, scalars, lr=None):
        """All-reduce (data parallel) and apply one Adam step; clears the gradient buffer."""
        if self.peer is not None:
            torch = _torch()
            ev = getattr(self, "allreduce_events", None)
            if ev is not None:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(torch.cuda.current_stream(self.m.device))
            self._sym_scalars.copy_(scalars)
            self.t += 1
            a, b = self._peer_args, self.peer
            cabi.call("htcn_peer_allreduce_adam", a["grads"][0], a["reduced"][0], a["scalars"][0], self.world, b.rank,
                      self.params.data_ptr(), self.adam_m.data_ptr(), self.adam_v.data_ptr(), self.n_flat, self.lr_t(lr),
                      self.beta1, self.beta2, self.eps, self._scalars_out.data_ptr(), a["f_in"][0], a["f_mid"][0],
                      b.own + 0 * 8 * 4, b.own + 1 * 8 * 4, b.own + 256, b.own + 256 + 64, self.t, self.m.stream_ptr())
            if ev is not None:
                e1.record(torch.cuda.current_stream(self.m.device))
                ev.append((e0, e1))
            if self.bf16:
                self._refresh_bf16_tables()
            return self._scalars_out.clone()
        if self.world > 1:
            ev = getattr(self, "allreduce_events", None)    # bench.py: [] -> (start, end) CUDA events of every exchange
            if ev is not None:
                torch = _torch()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(torch.cuda.current_stream(self.m.device))
            self.dist.all_reduce(self.grads)
            from .dist import allreduce_scalars
            scalars = allreduce_scalars(scalars, self.dist, self.world)
            if ev is not None:
                e1.record(torch.cuda.current_stream(self.m.device))
                ev.append((e0, e1))
        self.t += 1
        cabi.call("htcn_adam_step", self.params.data_ptr(), self.grads.data_ptr(), self.adam_m.data_ptr(),
                  self.adam_v.data_ptr(), self.n_flat, self.lr_t(lr), self.beta1, self.beta2, self.eps,
                  scalars[6:7].data_ptr(), 1, self.m.stream_ptr())
        if self.bf16:
            self._refresh_bf16_tables()
        return scalars

    def check_peer(self):
        """synchronises; raises if a flag wait of the fused all-reduce + Adam kernel ran into its spin limit"""
        if self.peer is not None and int(self.peer.view(256 + 64, (1,), _torch().int32).item()):
            raise RuntimeError("peer all-reduce: a rank did not arrive within the spin limit")

    def close(self):
        """unmap / free the symmetric gradient buffer (collective-free; call on every rank once training is over)"""
        if self.peer is not None:
            _torch().cuda.synchronize(self.m.device)
            self.grads = self.grads.clone()
            self.peer.close()
            self.peer = None

    def train_step(self, x_list, y_list, mask_list, state=None, lr=None, metrics=False, state_on_device=False,
                   mask_warmstart=None, x_gap=None, dropout_masks=None, neg_ids=None, loss_kind=None):
        """One optimisation step on one batch.  Returns dict(loss, user_count, n_valid, state [+ metrics]); ``state`` is
        the carried user state for the next batch (numpy, or the device tensor with ``state_on_device``)."""
        r = self.forward_backward(x_list, y_list, mask_list, state, metrics=metrics, mask_warmstart=mask_warmstart, x_gap=x_gap,
                                  dropout_masks=dropout_masks, neg_ids=neg_ids, loss_kind=loss_kind)
        sc = self.apply_gradients(r["scalars"], lr).cpu().numpy()
        out = dict(loss=float(sc[0]), user_count=float(sc[6]), n_valid=float(sc[7]))
        if metrics:
            out.update(recall1=float(sc[1]), recall5=float(sc[2]), recall10=float(sc[3]), mrr=float(sc[4]), mrp=float(sc[5]))
        if state_on_device:
            out["state"] = r["state"]
        else:
            from .weights import unpad_state
            out["state"] = unpad_state(r["state"].cpu().numpy(), self.m.layout_meta)
        return out

    # ------------------------------------------------------------------ inspection / checkpoint
    def named_gradients(self, user_count, clear=True):
        """Gradients of the reference loss (model.py:117) keyed by TF variable name, as fp64 numpy; call after
        ``forward_backward`` with user_count = its scalars[6]."""
        g = {k: (v.detach().cpu().numpy().astype(np.float64) / float(user_count)) for k, v in self.g.items()}
        if clear:
            self.grads.zero_()
        return self._to_tf_names(g)

    def state_dict(self):
        """Current weights keyed by the TF variable names of SURVEY.md A.6 (numpy fp32)."""
        return {k: v.astype(np.float32) for k, v in self._to_tf_names({k: v.detach().cpu().numpy() for k, v in self.p.items()}).items()}

    # checkpoint: the variables under their TF names plus what the reference's tf.train.Saver leaves out and a resumed run
    # needs to continue the same trajectory (Adam slots ARE in the Saver; the carried user state, the loader
    # cursor and the decayed learning rate are not)
    def save_checkpoint(self, path, state=None, loader=None, epoch=0):
        import json
        torch = _torch()
        torch.cuda.synchronize(self.m.device)
        out = {"w|" + k.replace("/", "|"): v for k, v in self.state_dict().items()}
        for tag, flat in (("Adam", self.adam_m), ("Adam_1", self.adam_v)):       # TF slot names
            views = {n: self._view(flat, n, sh).detach().cpu().numpy() for n, sh in self.spec}
            for k, v in self._to_tf_names(views).items():
                out[tag + "|" + k.replace("/", "|")] = v
        out["meta"] = np.frombuffer(json.dumps(dict(step=self.t, lr=self.lr, beta1=self.beta1, beta2=self.beta2, eps=self.eps,
                                                    epoch=int(epoch), precision=self.m.precision)).encode(), dtype=np.uint8)
        if state is not None:
            if hasattr(state, "detach"):        # device tensor = the padded device layout [B, G*128]: store [B, G*H]
                from .weights import unpad_state
                out["carried_state"] = unpad_state(state.detach().cpu().numpy(), self.m.layout_meta)
            else:
                out["carried_state"] = np.asarray(state, np.float32)
        if loader is not None and hasattr(loader, "state_dict"):
            out["loader"] = np.frombuffer(json.dumps(loader.state_dict()).encode(), dtype=np.uint8)
        np.savez(path, **out)

    def load_checkpoint(self, path, loader=None):
        """Restores weights, Adam slots, step counter and lr; returns dict(epoch, state) (state = carried user state or None)."""
        import json
        torch = _torch()
        z = np.load(path)
        meta = json.loads(bytes(z["meta"]).decode())

        def scatter(prefix, flat):
            from .weights import to_device_layout
            t = {k[len(prefix):].replace("|", "/"): z[k] for k in z.files if k.startswith(prefix)}
            named, _ = to_device_layout(t, "hier")       # TF shapes -> the 128-padded device layout
            named["wt"] = np.ascontiguousarray(named.pop("w_out").T)
            for n, sh in self.spec:
                self._view(flat, n, sh).copy_(torch.from_numpy(np.ascontiguousarray(named[n], dtype=np.float32)).reshape(sh))

        scatter("w|", self.params)
        scatter("Adam|", self.adam_m)
        scatter("Adam_1|", self.adam_v)
        self.grads.zero_()
        self.t, self.lr = int(meta["step"]), float(meta["lr"])
        if self.bf16:
            self._refresh_bf16_tables()
        if loader is not None and "loader" in z.files and hasattr(loader, "load_state_dict"):
            loader.load_state_dict(json.loads(bytes(z["loader"]).decode()))
        torch.cuda.synchronize(self.m.device)
        return dict(epoch=int(meta["epoch"]), state=z["carried_state"] if "carried_state" in z.files else None)

    def _to_tf_names(self, t):
        """flat-buffer views (device layout, every width padded to 128) -> the TF names / shapes of SURVEY A.6"""
        from .weights import from_device_layout
        return from_device_layout(t, self.m.layout_meta)


def lr_for_epoch(base_lr, epoch, lr_schedule=True):
    """run_hier_xing.py:265-269: the learning rate is divided by 5 when epoch 60 starts and again at epoch 120."""
    if not lr_schedule:
        return base_lr
    return base_lr / (5.0 if epoch >= 60 else 1.0) / (5.0 if epoch >= 120 else 1.0)


def run_hier(trainer: HierTCNTrainer, loader_train, epochs, epoch_batches_train, lr_schedule=True, start_epoch=0,
             on_epoch_end=None, verbose=False):
    """The training half of reference run_hier_xing.py:257-307: ``epoch_batches_train`` steps per epoch, the user state
    carried across batches (on the device here), the step-wise lr schedule.  Returns the per-epoch mean training loss."""
    state, hist = None, []
    for epoch in range(start_epoch, epochs):
        lr = lr_for_epoch(trainer.lr, epoch, lr_schedule)
        tot = 0.0
        for _ in range(epoch_batches_train):
            x_list, y_list, mask_list, _info = loader_train.get_batch()
            out = trainer.train_step(x_list, y_list, mask_list, state, lr=lr, state_on_device=True)
            state = out["state"]
            tot += out["loss"]
        hist.append(tot / max(1, epoch_batches_train))
        if verbose:
            print("epoch %d lr %.5f training loss %.5f" % (epoch, lr, hist[-1]))
        if on_epoch_end is not None:
            on_epoch_end(epoch, trainer)
    return hist
