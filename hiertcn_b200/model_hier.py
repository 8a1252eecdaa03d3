"""Host-side mirror of the reference's model surface for the hot path.

    reference                                   here
    ---------------------------------------     -------------------------------------------------
    model_hier(args,x,y,mask,state,...)          model_hier(...)  -> (CatalogScores, state)
      (model_hier.py:21-94)                      HierTCN.build / forward / loss / step
    model_tcn(args,x,...) (model_tcn.py:13-44)   model_tcn(...)
    sess.run([loss,state,ranks_float,...])       HierTCN.step(x_list,y_list,mask_list,state)
      (run_hier_xing.py:145-149,301-302)

PyTorch is used for device memory, streams and pinned staging only; all arithmetic is done by the
sm_100a kernels behind the C ABI (hiertcn_b200/_cabi.py -> libhtcn.so).  There is no CPU path: without
the library or without a CUDA device every entry point raises.

Execution order (the GRU input is the teacher-forced mean of the session's true items,
so the S-step recurrence is hoisted out of the session loop):

    K1 gather+meanpool -> K3 GRU over sessions (+ state half of the in-projection)
       -> K2 in-projection + causal conv stack -> K4 catalog scoring with fused CE / rank / top-k
"""
from __future__ import annotations

import math
import os

import numpy as np

from . import _cabi as cabi
from .weights import hier_weight_shapes, init_weights

D = 128


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise cabi.HtcnError("hiertcn_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch


class CatalogScores:
    """Lazy stand-in for the reference's ``pred_all [B,T,N]`` (model_hier.py:76-79).  The logits are
    never written to HBM; losses / ranks / top-k are produced by streaming reductions on demand.
    ``materialize()`` builds the dense tensor for small catalogs (tests, debugging)."""

    def __init__(self, model, hout, Q, row_of, y_rows, y_id, B, T):
        self.model, self.hout, self.Q = model, hout, Q
        self.row_of, self.y_rows, self.y_id, self.B, self.T = row_of, y_rows, y_id, B, T
        self._cache = {}
        self.generation = model.generation

    def check_fresh(self):
        """``hout`` / ``row_of`` and the cached loss rows are views of the model's reusable workspaces: a later
        ``forward`` / ``step`` / training step on the same model overwrites them.  Using a stale handle raises instead of
        returning the other batch's numbers."""
        if self.generation != self.model.generation:
            raise cabi.HtcnError("stale CatalogScores: the model ran another forward since this handle was created "
                                 "(its workspaces were reused); score / materialize it before the next forward")

    @property
    def shape(self):
        return (self.B, self.T, self.model.N)

    def materialize(self):
        """[B,T,N] fp32 numpy; masked positions are zero rows like ``pred *= mask_y`` (model.py:105)."""
        torch = _torch()
        self.check_fresh()
        m = self.model
        out = np.zeros((self.B * self.T, m.N), dtype=np.float32)
        if self.Q:
            lg = torch.empty((self.Q, m.N), dtype=torch.float32, device=m.device)
            cabi.call("htcn_score_logits", self.hout.data_ptr(), m.k4_dtype, self.Q, m.wt.data_ptr(), m.act_dtype,
                      m.b_out.data_ptr(), m.N, lg.data_ptr(), m.stream_ptr())
            if m.l2_normalize:          # model_tcn.py:42-43
                cabi.call("htcn_scale_rows", lg.data_ptr(), m.row_scale(self).data_ptr(), self.Q, m.N, m.stream_ptr())
            rows = self.row_of.cpu().numpy()
            out[rows >= 0] = lg.cpu().numpy()[rows[rows >= 0]]
        return out.reshape(self.B, self.T, m.N)


class HierTCN:
    """HierTCN forward + catalog scoring on one B200.

    args: namespace with the reference's hyper-parameter names (hiertcn_b200.args.make_args()).
    weights: dict keyed by the TF variable names of SURVEY.md A.6 (hiertcn_b200.weights); random init if None.
    """

    def __init__(self, args, weights=None, device=None, precision=None, seed=1234):
        self.args = args
        self.precision = precision or getattr(args, "precision", "bf16")
        if self.precision not in ("f32", "bf16"):
            raise ValueError("precision must be 'f32' or 'bf16'")
        if args.model_type != "hier" or args.model_low_type != "tcn":
            raise NotImplementedError("only model_type='hier' with model_low_type='tcn' is the hot path")
        # widths below 128 run zero-padded to 128 (hiertcn_b200.weights.to_device_layout); levels that change the width get
        # the 1x1 down-sample residual of customized_tcn_cell.py:102-106
        if max(int(args.hidden_dim), int(getattr(args, "emb_dim", 128))) > 128 or max(args.tcn_channel) > 256:
            raise NotImplementedError("hidden_dim / emb_dim above 128 or tcn_channel above 256: the sm_100a kernels run "
                                      "128-wide blocks")
        if max(args.tcn_channel) > 128 and self.precision != "f32":
            # 129..256-channel levels (the single-level default of args.py:310-311) run as two 128-wide planes on the fp32
            # kernels; the tcgen05 tier is built for levels up to 128 channels
            raise NotImplementedError("tcn_channel above 128 needs precision='f32' (two-plane fp32 kernels)")
        for flag in ("has_batchnorm", "has_layernorm", "has_impression"):
            if getattr(args, flag, False):
                raise NotImplementedError("%s is outside the hot path (SURVEY.md A.8)" % flag)
        if getattr(args, "has_gap", False) and getattr(args, "train_gap", False):
            raise NotImplementedError("train_gap (learned gap bandwidth, model_hier.py:42-45); the fixed-bandwidth decay is built")
        self.l2_normalize = bool(getattr(args, "l2_normalize", False))     # model_tcn.py:42-43
        # bf16 tier: the CE (+ rank) sweep with the softmax exponent folded into the tensor-core product
        # (htcn_score_ce_rank_folded).  HTCN_K4_FOLD=0 (or an explicit HTCN_K4_EPI / HTCN_K4_CTA_GROUP=1) selects the older
        # epilogues of htcn_score_ce_rank_topk for A/B runs.
        self.k4_fold_enabled = True
        # args.dropout > 0 acts in HierTCNTrainer only (tf.layers.Dropout is the identity when training=False)
        self.N = int(args.item_num)
        self.G = int(args.num_layer)
        self.K = int(args.kernel_size)
        self.n_levels = len(args.tcn_channel)
        self.host_weights = weights
        self.seed = seed
        self.device = device
        self.built = False
        self._init_runtime_state()

    def _init_runtime_state(self):
        """workspaces, pinned staging sets (two input sets + three result sets, see ``stage`` / ``step_async``) and the
        generation counter that invalidates ``CatalogScores`` handles whose workspaces were reused"""
        self._ws = {}
        self._pin, self._pin_ev, self._pin_next, self._res_next = {}, [None, None], 0, 0
        self.generation = 0

    # ------------------------------------------------------------------ build
    def build(self):
        torch = _torch()
        cabi.load()
        if self.device is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        if not cabi.load().htcn_device_ok():
            raise cabi.HtcnError("current CUDA device is not compute capability 10.x (B200)")
        a = self.args
        w = self.host_weights
        if w is None:
            w = init_weights(hier_weight_shapes(self.N, a.hidden_dim, self.G, tuple(a.tcn_channel), self.K,
                                                getattr(a, "emb_dim", 128)), seed=self.seed)
        from .weights import fold_weightnorm, to_device_layout
        w = fold_weightnorm(w)
        lay, meta = to_device_layout(w, "hier")          # every width zero-padded to 128
        self.layout_meta = meta
        if meta["G"] != self.G or meta["K"] != self.K or len(meta["channels"]) != self.n_levels:
            raise ValueError("weights do not match args: num_layer %d/%d kernel_size %d/%d levels %d/%d"
                             % (meta["G"], self.G, meta["K"], self.K, len(meta["channels"]), self.n_levels))
        ed = meta["ed"]
        dev = self.device

        def up(x):
            return torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).to(dev)

        # emb_dim < 128: the table is stored PACKED (emb_pitch floats per row, e.g. 100 -> 400 B rows) and K1 zero-fills
        # the output rows beyond it (same math as a zero-padded table, 22% fewer gathered bytes at config 2's 100-d)
        self.emb_pitch = min(D, -(-ed // 4) * 4)
        self.E, self.b_emb = up(lay["E"][:, :self.emb_pitch]), up(lay["b_emb"])
        self.w_in_x, self.w_in_state = up(lay["w_in_x"]), up(lay["w_in_state"])
        self.conv_w = [up(lay[f"conv_w{l}"]) for l in range(self.n_levels)]
        self.conv_b = [up(lay[f"conv_b{l}"]) for l in range(self.n_levels)]
        self.ds_w = [up(lay[f"ds_w{l}"]) if meta["ds"][l] else None for l in range(self.n_levels)]
        self.ds_b = [up(lay[f"ds_b{l}"]) if meta["ds"][l] else None for l in range(self.n_levels)]
        self.gru = [tuple(up(lay[f"{n}{g}"]) for n in ("gate_w", "gate_b", "cand_w", "cand_b")) for g in range(self.G)]
        self.b_out = up(lay["b_out"])
        self._w_out_host = lay["w_out"]                 # fp32 master of the output table (hiertcn_b200.train, bf16 tier)
        w_out = up(lay["w_out"])                        # [128, N] TF layout (rows beyond the last level's width are zero)
        self._finish_build_head(w_out, meta)
        self.wt_f32 = self.wt if self.precision == "f32" else None
        torch.cuda.synchronize(dev)
        del w_out
        self.refresh_pointer_tables()
        self.built = True
        return self

    def _finish_build_head(self, w_out, meta):
        """dtypes + the scoring table W_out^T; ``wide``: some level has 129..256 channels (two 128-wide planes, fp32 kernels)"""
        torch = _torch()
        dev = self.device
        self.act_dtype = cabi.HTCN_BF16 if self.precision == "bf16" else cabi.HTCN_F32
        tdt = torch.bfloat16 if self.precision == "bf16" else torch.float32
        self.act_torch_dtype = tdt
        self.wide = bool(meta["wide"])
        self.level_planes = list(meta["planes"])
        self.head_planes = self.level_planes[-1] if self.level_planes else 1
        if self.wide and self.precision != "f32":
            raise NotImplementedError("tcn_channel above 128 needs precision='f32'")
        # precision / dtype code of the K4 calls: 256-wide user embeddings are block-planar [2][Q][128]
        self.k4_dtype = cabi.HTCN_F32_W256 if self.head_planes == 2 else self.act_dtype
        self.n_out = int(w_out.shape[1])               # == N, or this rank's catalog shard (hiertcn_b200.dist)
        pitch = cabi.WT_PITCH_BF16 if self.precision == "bf16" else D * self.head_planes   # bf16 rows carry the bias
        self.wt = torch.empty((self.n_out, pitch), dtype=tdt, device=dev)  # W_out^T, K-major rows
        cabi.call("htcn_prepare_wout", w_out.data_ptr(), self.b_out.data_ptr(), self.n_out, self.wt.data_ptr(),
                  self.k4_dtype, self.stream_ptr())

    def refresh_pointer_tables(self):
        """host arrays of device pointers the C ABI takes (rebuilt when the trainer re-homes the parameters)"""
        self._conv_w_pp = cabi.ptr_array([t.data_ptr() for t in self.conv_w])
        self._conv_b_pp = cabi.ptr_array([t.data_ptr() for t in self.conv_b])
        self.has_ds = any(t is not None for t in self.ds_w)
        self._ds_w_pp = cabi.ptr_array([t.data_ptr() if t is not None else 0 for t in self.ds_w]) if self.has_ds else (None, None)
        self._ds_b_pp = cabi.ptr_array([t.data_ptr() if t is not None else 0 for t in self.ds_b]) if self.has_ds else (None, None)
        if self.G:
            self._gru_pp = [cabi.ptr_array([l[i].data_ptr() for l in self.gru]) for i in range(4)]

    def stream_ptr(self):
        return _torch().cuda.current_stream(self.device).cuda_stream

    def _buf(self, name, shape, dtype):
        torch = _torch()
        t = self._ws.get(name)
        n = int(np.prod(shape))
        if t is None or t.dtype != dtype or t.numel() < n:
            t = torch.empty(max(n, 1), dtype=dtype, device=self.device)
            self._ws[name] = t
        return t[:n].view(*shape) if n else t[:0]

    # ------------------------------------------------------------------ host -> device staging
    def _pinned(self, slot, name, shape, dtype):
        """reusable pinned host staging tensor (grow-only) of staging set ``slot``; returns (torch tensor, numpy view)"""
        torch = _torch()
        key = (slot, name)
        n = int(np.prod(shape))
        t = self._pin.get(key)
        if t is None or t.dtype != dtype or t.numel() < n:
            t = torch.empty(max(n, 1), dtype=dtype).pin_memory()
            self._pin[key] = t
        v = t[:n].view(*shape) if n else t[:0]
        return v, v.numpy()

    def stage(self, x_list, y_list, mask_list, state=None, neg_ids=None, mask_warmstart=None, x_gap=None):
        """Pack the reference batch layout (data_loader.dequeue) straight into pinned staging buffers and copy it to
        the device (asynchronously, on the current stream).  Two staging sets alternate, each guarded by an event,
        so the host can prepare batch i+1 while the copies of batch i are still in flight.
        ``mask_warmstart [B,T]`` (0/1): multiplied into mask_y like model.py:102-103 -- masked positions are not scored
        and do not count in the per-user means (they still feed the GRU's mean-pool, as in the reference).
        ``x_gap``: S-list of [B,1] time gaps; with ``args.has_gap`` the carried state is decayed by
        exp(-gap / args.gap_bandwidth) before every slot (model_hier.py:40-47).  The decay of slot s+1 is folded into the
        reset mask of slot s (both multiply the state between two GRU steps) and that of slot 0 into the incoming state.
        Returns a dict of device tensors + host metadata; H2D bytes are in ['h2d_bytes']."""
        torch = _torch()
        if not self.built:
            self.build()
        slot = self._pin_next
        self._pin_next ^= 1
        if self._pin_ev[slot] is not None:
            self._pin_ev[slot].synchronize()                    # the previous copies out of this set have completed
        S = len(x_list)
        lens = [int(np.shape(x)[1]) for x in x_list]
        B = int(np.shape(x_list[0])[0])
        T = int(sum(lens))
        slot_off = np.zeros(S + 1, dtype=np.int32)
        slot_off[1:] = np.cumsum(lens)
        i32, f32 = torch.int32, torch.float32
        tx, nx = self._pinned(slot, "x_id", (B, T), i32)
        ty, ny = self._pinned(slot, "y_id", (B, T), i32)
        tm, nm = self._pinned(slot, "mask", (S, B), f32)
        for s in range(S):                                      # cast + concat in one pass, no temporaries
            nx[:, slot_off[s]:slot_off[s + 1]] = x_list[s]
            ny[:, slot_off[s]:slot_off[s + 1]] = y_list[s]
            nm[s] = np.asarray(mask_list[s]).reshape(-1)
        decay0 = None
        if x_gap is not None and getattr(self.args, "has_gap", False):
            bw = np.float32(self.args.gap_bandwidth)
            dec = [np.exp(-np.asarray(g, np.float32).reshape(-1) / bw) for g in x_gap]
            decay0 = dec[0]
            for s in range(S - 1):
                nm[s] *= dec[s + 1]
        valid = ny.reshape(-1) > 0
        host_extra = {}
        if mask_warmstart is not None:
            warm = np.asarray(mask_warmstart).reshape(-1) > 0
            valid &= warm
            tl, nl = self._pinned(slot, "y_loss", (B, T), i32)       # y_id with the warm-start mask applied: the loss side
            nl[:] = ny * warm.reshape(B, T)
            host_extra["y_loss"] = tl
        tr, nr = self._pinned(slot, "row_of", (B * T,), i32)
        np.cumsum(valid, dtype=np.int32, out=nr)
        Q = int(nr[-1]) if nr.size else 0
        nr -= 1
        nr[~valid] = -1
        tq, nq = self._pinned(slot, "y_rows", (Q,), i32)
        if Q:
            nq[:] = ny.reshape(-1)[valid]
        host = dict(x_id=tx, y_id=ty, mask=tm, row_of=tr, y_rows=tq, **host_extra)
        dev, nbytes = {}, 0
        if hasattr(state, "data_ptr"):
            # device-resident carried state: the previous step's state_out stays in HBM instead of
            # round-tripping through host numpy every batch like the reference does (run_hier_xing.py:291,301)
            dev["state"] = state.to(device=self.device, dtype=f32).contiguous()
            if decay0 is not None:
                dev["state"] = dev["state"].clone()
                d0 = torch.from_numpy(np.ascontiguousarray(decay0)).to(self.device)
                cabi.call("htcn_scale_rows", dev["state"].data_ptr(), d0.data_ptr(), B, self.G * 128, self.stream_ptr())
        else:
            ts, ns_ = self._pinned(slot, "state", (B, self.G * 128), f32)
            if state is None:
                ns_[:] = 0.0
            else:
                from .weights import pad_state
                ns_[:] = pad_state(state, self.layout_meta)       # hidden_dim < 128: zero-padded units
                if decay0 is not None:
                    ns_ *= decay0[:, None]
            host["state"] = ts
        if neg_ids is not None and not hasattr(neg_ids, "data_ptr"):
            tn, nn = self._pinned(slot, "neg_ids", tuple(np.shape(neg_ids)), i32)
            nn[:] = neg_ids
            host["neg_ids"] = tn
        # H2D on a dedicated copy stream: the copies of batch i+1 overlap the kernels of batch i
        if not hasattr(self, "_copy_stream"):
            self._copy_stream = torch.cuda.Stream(device=self.device)
        compute = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(self._copy_stream):
            for k, t in host.items():
                dev[k] = t.to(self.device, non_blocking=True)
                nbytes += t.numel() * t.element_size()
        ev = torch.cuda.Event()
        ev.record(self._copy_stream)
        compute.wait_event(ev)
        for k in host:
            dev[k].record_stream(compute)
        self._pin_ev[slot] = ev
        dev.update(B=B, T=T, S=S, Q=Q, slot_off=slot_off, h2d_bytes=nbytes)
        return dev

    # ------------------------------------------------------------------ forward (K1 -> K3 -> K2)
    def forward(self, x_list=None, y_list=None, mask_list=None, state=None, staged=None, mask_warmstart=None, x_gap=None):
        """Hierarchical forward up to the user embeddings.  Returns (CatalogScores, state_out [B,G*H] device)."""
        if not self.built:
            self.build()
        torch = _torch()
        d = staged if staged is not None else self.stage(x_list, y_list, mask_list, state, None, mask_warmstart, x_gap)
        self.generation += 1            # handles of earlier forwards now point at overwritten workspaces
        B, T, S, Q = d["B"], d["T"], d["S"], d["Q"]
        st = self.stream_ptr()
        f32 = torch.float32
        slot_p, slot_keep = cabi.int_array(d["slot_off"])
        xe = self._buf("xe", (B * T, D), self.act_torch_dtype)
        yp = self._buf("yp", (S, B, D), f32)
        cabi.call("htcn_gather_meanpool", self.E.data_ptr(), self.emb_pitch, self.b_emb.data_ptr(), self.N, d["x_id"].data_ptr(),
                  d["y_id"].data_ptr(), slot_p, B, T, S, xe.data_ptr(), self.act_dtype, yp.data_ptr(), st)
        sbias = self._buf("sbias", (S, B, D), f32)
        state_out = torch.empty((B, self.G * 128), dtype=f32, device=self.device)
        # bf16 tier: tensor-core GRU (bf16 operands, fp32 state) when the stack has 2 layers; set self.k3_tcgen05 = False
        # to keep the fp32 FFMA recurrence
        k3_bf16 = self.precision == "bf16" and self.G == 2 and getattr(self, "k3_tcgen05", True)
        k3_scratch = self._buf("k3_scratch", (cabi.gru_scratch_bytes(B) // 4,), f32) if k3_bf16 else None
        cabi.call("htcn_gru_sessions", yp.data_ptr(), d["mask"].data_ptr(), d["state"].data_ptr(),
                  self._gru_pp[0][0], self._gru_pp[1][0], self._gru_pp[2][0], self._gru_pp[3][0], self.G,
                  self.w_in_state.data_ptr(), B, S, cabi.HTCN_BF16 if k3_bf16 else cabi.HTCN_F32,
                  k3_scratch.data_ptr() if k3_bf16 else None, None, sbias.data_ptr(), state_out.data_ptr(), st)
        hout = self._buf("hout", (self.head_planes * max(Q, 1), D), self.act_torch_dtype)   # wide head: planes [P][Q][128]
        k2_precision = self._k2_precision()
        if self.wide:
            scratch = self._buf("k2_scratch", (6 * B * T, D), f32)
            planes_p, planes_keep = cabi.int_array(self.level_planes)
            cabi.call("htcn_tcn_forward_wide", xe.data_ptr(), self.act_dtype, self.w_in_x.data_ptr(), sbias.data_ptr(),
                      self._conv_w_pp[0], self._conv_b_pp[0], self._ds_w_pp[0], self._ds_b_pp[0], planes_p, self.n_levels,
                      self.K, slot_p, B, T, S, d["row_of"].data_ptr(), hout.data_ptr(), max(Q, 1), scratch.data_ptr(), st)
            cabi.note_launches(self.n_levels + 1 + sum(t is not None for t in self.ds_w))
            del slot_keep, planes_keep
            return CatalogScores(self, hout, Q, d["row_of"], d["y_rows"], d.get("y_loss", d["y_id"]), B, T), state_out
        if k2_precision == cabi.HTCN_F32:
            scratch = self._buf("k2_scratch", ((3 if self.has_ds else 2) * B * T, D), f32)
        else:       # bf16 weight tiles + pointer table + biases of the fused tcgen05 conv stack
            scratch = self._buf("k2_scratch_bf16", (cabi.tcn_scratch_floats(self.n_levels, self.K),), f32)
        cabi.call("htcn_tcn_forward", xe.data_ptr(), self.act_dtype, k2_precision, self.w_in_x.data_ptr(),
                  sbias.data_ptr(), self._conv_w_pp[0], self._conv_b_pp[0], self._ds_w_pp[0], self._ds_b_pp[0],
                  self.n_levels, self.K, slot_p, B, T, S,
                  d["row_of"].data_ptr(), hout.data_ptr(), self.act_dtype,
                  scratch.data_ptr() if scratch is not None else None, st)
        # fp32: one launch per level + the in-projection; bf16: weight-tile preparation + the fused stack
        cabi.note_launches(2 if k2_precision == cabi.HTCN_BF16 else self.n_levels + 1 + sum(t is not None for t in self.ds_w))
        if k3_bf16:
            cabi.note_launches(1)                   # k3_prepare_weights (htcn_gru_sessions counts one launch)
        del slot_keep
        scores = CatalogScores(self, hout, Q, d["row_of"], d["y_rows"], d.get("y_loss", d["y_id"]), B, T)
        return scores, state_out

    def _k2_precision(self):
        # bf16 tier: fused tcgen05 conv stack (k2_tcn_bf16.cu); set self.k2_tcgen05 = False to run the fp32 FFMA
        # conv stack on bf16 inputs/outputs instead (more accurate, slower)
        return cabi.HTCN_BF16 if (self.precision == "bf16" and getattr(self, "k2_tcgen05", True)) else cabi.HTCN_F32

    # ------------------------------------------------------------------ loss / metrics / top-k (K4)
    def n_split_for(self, Q, n_items):
        forced = getattr(self, "force_n_split", 0)
        if forced:
            return int(max(1, min(forced, max(1, n_items // 256))))
        from .dist import choose_n_split
        if getattr(self, "_sms", None) is None:
            self._sms = int(_torch().cuda.get_device_properties(self.device).multi_processor_count)
        return choose_n_split(Q, n_items, self._sms)

    @property
    def k4_fold(self):
        import os
        return (getattr(self, "k4_fold_enabled", True) and os.environ.get("HTCN_K4_FOLD", "1") != "0" and "HTCN_K4_EPI" not in os.environ
                and os.environ.get("HTCN_K4_CTA_GROUP", "2") == "2")

    def score(self, scores: CatalogScores, ce=True, rank=True, topk=0):
        """One streaming sweep over the catalog.  Returns dict of device tensors:
        loss_row [Q], rank_row [Q] and, with topk, topk_val / topk_idx [Q,k]."""
        torch = _torch()
        scores.check_fresh()
        if self.n_out != self.N:
            raise cabi.HtcnError("this model holds a catalog shard (%d of %d output rows): score it through "
                                 "hiertcn_b200.dist.ShardedCatalogScorer (or HierTCN.topk with the shard offset)"
                                 % (self.n_out, self.N))
        Q = scores.Q
        key = (ce, rank, topk)
        if key in scores._cache:
            return scores._cache[key]
        out = {}
        if Q == 0:
            scores._cache[key] = out
            return out
        st = self.stream_ptr()
        f32, i32 = torch.float32, torch.int32
        flags = (cabi.SCORE_CE if ce else 0) | (cabi.SCORE_RANK if rank else 0)
        ns = self.n_split_for(Q, self.N)
        zy = self._buf("zy", (Q,), f32)
        pm = self._buf("pm", (ns, Q), f32) if ce else None
        ps = self._buf("ps", (ns, Q), f32) if ce else None
        pc = self._buf("pc", (ns, Q), i32) if rank else None
        tv = ti = None
        P = lambda t: t.data_ptr() if t is not None else None  # noqa: E731
        need_t = ce or rank
        if need_t:      # target logits first (own launch so the sweep can be timed on its own)
            cabi.call("htcn_target_logit", scores.hout.data_ptr(), self.k4_dtype, Q, self.wt.data_ptr(),
                      self.b_out.data_ptr(), self.N, 0, scores.y_rows.data_ptr(), zy.data_ptr(), st)
        ev = getattr(self, "sweep_events", None)
        if ev is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(torch.cuda.current_stream(self.device))
        fused = bool(topk) and ce and rank and not self.l2_normalize    # loss + rank + top-k: two catalog sweeps instead of three
        zy_fin = zy
        if self.l2_normalize and self.head_planes > 1:
            raise NotImplementedError("l2_normalize with a 256-channel last level")
        if self.l2_normalize and flags:
            # l2-normalised head (model_tcn.py:42-43): CE on row_scale * z with row_scale = 1/||z|| from the catalog's Gram
            # matrix; ranks are invariant under the positive scale
            rs = self.row_scale(scores)
            pm = pm if pm is not None else self._buf("pm", (ns, Q), f32)
            ps = ps if ps is not None else self._buf("ps", (ns, Q), f32)
            pc = pc if pc is not None else self._buf("pc", (ns, Q), i32)
            zy_fin = self._buf("zy_scaled", (Q,), f32)
            cabi.call("htcn_score_ce_rank_l2norm", scores.hout.data_ptr(), self.act_dtype, Q, self.wt.data_ptr(),
                      self.b_out.data_ptr(), self.N, 0, scores.y_rows.data_ptr(), zy.data_ptr(), rs.data_ptr(), ns,
                      pm.data_ptr(), ps.data_ptr(), pc.data_ptr(), zy_fin.data_ptr(), st)
        elif fused:
            nbytes = int(cabi.load().htcn_topk_workspace_bytes(self.k4_dtype, Q, self.N, topk, ns))
            ws = self._buf("topk_ws", (nbytes,), torch.uint8)
            ov = torch.empty((Q, topk), dtype=f32, device=self.device)
            oi = torch.empty((Q, topk), dtype=i32, device=self.device)
            ovf = torch.zeros(1, dtype=i32, device=self.device)
            cabi.call("htcn_score_ce_rank_topk_fused", scores.hout.data_ptr(), self.k4_dtype, Q, self.wt.data_ptr(),
                      self.b_out.data_ptr(), self.N, 0, scores.y_rows.data_ptr(), zy.data_ptr(), topk, ns, ws.data_ptr(),
                      nbytes, P(pm), P(ps), P(pc), ov.data_ptr(), oi.data_ptr(), ovf.data_ptr(), st)
            out.update(topk_val=ov, topk_idx=oi)
            self._topk_overflow = ovf
        elif ce and self.precision == "bf16" and self.k4_fold:      # (read per call: the sweep scripts switch variants)
            # the default loss sweep of the bf16 tier: exponent and rank compare folded into the tensor-core product
            ws = self._buf("fold_ws", (cabi.score_fold_ws_bytes(Q),), torch.uint8)
            cabi.call("htcn_score_ce_rank_folded", scores.hout.data_ptr(), Q, self.wt.data_ptr(), self.N, 0,
                      scores.y_rows.data_ptr(), zy.data_ptr(), flags, ns, P(pm), P(ps), P(pc), ws.data_ptr(), st)
        elif flags:
            cabi.call("htcn_score_ce_rank_topk", scores.hout.data_ptr(), self.k4_dtype, Q, self.wt.data_ptr(),
                      self.b_out.data_ptr(), self.N, 0, scores.y_rows.data_ptr(), zy.data_ptr(), 1,
                      flags, 0, ns, P(pm), P(ps), P(pc), None, None, st)
            cabi.note_launches(-1)      # have_target=1: the sweep call launched one kernel, not two
        if ev is not None:
            e1.record(torch.cuda.current_stream(self.device))
            ev.append((e0, e1, 2.0 * Q * 128 * self.N))
        if ce or rank:
            loss_row = self._buf("loss_row", (Q,), f32) if ce else None
            rank_row = self._buf("rank_row", (Q,), f32) if rank else None
            cabi.call("htcn_score_finish", P(pm) if ce else None, P(ps) if ce else None, P(pc) if rank else None, ns, Q,
                      scores.y_rows.data_ptr(), zy_fin.data_ptr(), P(loss_row), P(rank_row), st)
            if ce and self.precision == "bf16" and self.n_out == self.N and not self.l2_normalize:
                # rows whose target is > 88 nats below the best logit overflow the target-referenced partial sum: redo them
                cabi.call("htcn_score_ce_repair", scores.hout.data_ptr(), self.k4_dtype, Q, self.wt.data_ptr(), self.N,
                          zy.data_ptr(), loss_row.data_ptr(), None, st)
            out.update(loss_row=loss_row, rank_row=rank_row, target_logit=zy_fin)
        if fused and int(self._topk_overflow.item()):          # pathological ties overflowed a candidate list
            out.update(self.topk(scores.hout, Q, topk))
        elif topk and not fused:
            out.update(self.topk(scores.hout, Q, topk))
            if self.l2_normalize:       # same order, normalised values
                cabi.call("htcn_scale_rows", out["topk_val"].data_ptr(), self.row_scale(scores).data_ptr(), Q, topk, st)
        scores._cache[key] = out
        return out

    def row_scale(self, scores: CatalogScores):
        """[Q] 1 / max(||z_q||, 1e-6) of every scored row's logits over the whole catalog (tf.nn.l2_normalize, epsilon 1e-12
        on the squared norm) from the quadratic form h^T G h + 2 h.c + s; the catalog's Gram data (G, c, s) is built once
        per set of output weights (``invalidate_gram`` after an update)."""
        torch = _torch()
        if "row_scale" in scores._cache:
            return scores._cache["row_scale"]
        f32 = torch.float32
        if getattr(self, "_gram", None) is None:
            n_scr = int(cabi.load().htcn_catalog_gram_scratch_floats())
            scr = torch.empty(n_scr, dtype=f32, device=self.device)
            self._gram = torch.empty(D * D + D + 4, dtype=f32, device=self.device)
            cabi.call("htcn_catalog_gram", self.wt.data_ptr(), self.act_dtype, self.b_out.data_ptr(), self.n_out,
                      scr.data_ptr(), self._gram.data_ptr(), self.stream_ptr())
        rs = torch.empty(max(scores.Q, 1), dtype=f32, device=self.device)
        if scores.Q:
            cabi.call("htcn_logit_rownorm", scores.hout.data_ptr(), self.act_dtype, scores.Q, self._gram.data_ptr(), 1e-12,
                      rs.data_ptr(), self.stream_ptr())
        scores._cache["row_scale"] = rs
        return rs

    def invalidate_gram(self):
        self._gram = None

    def topk(self, hout, Q, k, n0=0):
        """Exact top-k of every row of ``hout`` over this model's catalog (shard).  Returns dict(topk_val, topk_idx)
        [Q,k] device tensors in tf.nn.top_k order.  Uses the two-pass tensor-core method on large bf16 catalogs and
        redoes the call with the heap sweep if a candidate list overflowed."""
        torch = _torch()
        f32, i32 = torch.float32, torch.int32
        tiles_q = max(1, math.ceil(Q / 128))
        ns = int(max(1, min(math.ceil(2 * 148 / tiles_q), 32, max(1, self.n_out // 256))))
        nbytes = int(cabi.load().htcn_topk_workspace_bytes(self.k4_dtype, Q, self.n_out, k, ns))
        ws = self._buf("topk_ws", (nbytes,), torch.uint8)
        ov = torch.empty((Q, k), dtype=f32, device=self.device)
        oi = torch.empty((Q, k), dtype=i32, device=self.device)
        ovf = torch.zeros(1, dtype=i32, device=self.device)
        cabi.call("htcn_score_topk", hout.data_ptr(), self.k4_dtype, Q, self.wt.data_ptr(), self.b_out.data_ptr(), self.n_out,
                  n0, k, ns, ws.data_ptr(), nbytes, ov.data_ptr(), oi.data_ptr(), ovf.data_ptr(), self.stream_ptr())
        if int(ovf.item()):                                   # pathological ties: exact heap path
            tv = torch.empty((ns, Q, k), dtype=f32, device=self.device)
            ti = torch.empty((ns, Q, k), dtype=i32, device=self.device)
            cabi.call("htcn_score_ce_rank_topk", hout.data_ptr(), self.k4_dtype, Q, self.wt.data_ptr(), self.b_out.data_ptr(),
                      self.n_out, n0, None, None, 1, cabi.SCORE_TOPK, k, ns, None, None, None, tv.data_ptr(), ti.data_ptr(),
                      self.stream_ptr())
            cabi.call("htcn_topk_merge", tv.data_ptr(), ti.data_ptr(), ns, Q, k, ov.data_ptr(), oi.data_ptr(), self.stream_ptr())
        return dict(topk_val=ov, topk_idx=oi)

    def loss(self, scores: CatalogScores, metrics=True, per_position=False):
        """Softmax-CE loss with the reference's two-level masked mean (model.py:105-117) and, with
        ``metrics``, calc_metric_fast (loss.py:163-221).  Returns dict of device tensors:
        scalars[8] = {loss, recall@1, recall@5, recall@10, mrr, mrp, user_count, n_valid} (+ [B,T] maps)."""
        torch = _torch()
        r = self.score(scores, ce=True, rank=metrics)
        B, T = scores.B, scores.T
        f32 = torch.float32
        scalars = torch.empty(8, dtype=f32, device=self.device)
        maps = {}
        if per_position:        # True = all three [B,T] maps; or an iterable of names (the reference's eval loop fetches
            names = ("loss_bt", "ranks", "ranks_float") if per_position is True else tuple(per_position)   # ranks_float only)
            for n in names:
                if n not in ("loss_bt", "ranks", "ranks_float"):
                    raise ValueError("per_position map %r" % (n,))
                maps[n] = torch.empty((B, T), dtype=f32, device=self.device)
        P = lambda t: t.data_ptr() if t is not None else None  # noqa: E731
        cabi.call("htcn_loss_metrics_reduce", P(r.get("loss_row")), P(r.get("rank_row")), scores.row_of.data_ptr(),
                  scores.y_id.data_ptr(), B, T, self.N, P(maps.get("loss_bt")), P(maps.get("ranks")),
                  P(maps.get("ranks_float")), self._buf("user_part", (B, 8), f32).data_ptr(), scalars.data_ptr(),
                  self.stream_ptr())
        return dict(scalars=scalars, **maps)

    def sampled_loss_mean(self, scores: CatalogScores, neg_ids, kind=None):
        """Sampled ranking loss reduced with the reference's two-level masked mean (model.py:111-117).
        Returns a device tensor scalars[8] whose element 0 is the loss."""
        torch = _torch()
        rows = self.sampled_loss(scores, neg_ids, kind)
        scalars = torch.empty(8, dtype=torch.float32, device=self.device)
        cabi.call("htcn_loss_metrics_reduce", rows.data_ptr(), None, scores.row_of.data_ptr(), scores.y_id.data_ptr(),
                  scores.B, scores.T, self.N, None, None, None,
                  self._buf("user_part", (scores.B, 8), torch.float32).data_ptr(), scalars.data_ptr(), self.stream_ptr())
        return scalars

    def sampled_loss(self, scores: CatalogScores, neg_ids, kind=None):
        """Sampled ranking loss (reference loss.py:22-71) of the user embeddings against the rows of the
        output table W_out^T: positive = the true next item, negatives = ``neg_ids [Q,k]`` (host or device)."""
        torch = _torch()
        scores.check_fresh()
        if self.n_out != self.N:
            raise cabi.HtcnError("sampled_loss gathers rows of the whole output table; this model holds a catalog shard")
        if self.head_planes > 1:
            raise NotImplementedError("sampled ranking losses with a 256-channel last level")
        a = self.args
        kind = kind or (a.loss if a.loss in cabi.LOSS_KINDS else "hinge_logsigmoid")
        if kind not in cabi.LOSS_KINDS:
            raise ValueError("sampled loss kind %r" % kind)
        neg = neg_ids if hasattr(neg_ids, "data_ptr") else torch.from_numpy(np.ascontiguousarray(neg_ids, np.int32)).to(self.device)
        Q, k = neg.shape
        assert Q == scores.Q
        out = torch.empty(Q, dtype=torch.float32, device=self.device)
        if (self.precision == "bf16" and self.wt_f32 is None and self.act_dtype == cabi.HTCN_BF16
                and os.environ.get("HTCN_SAMPLED_BF16TAB")):
            # opt-in (HTCN_SAMPLED_BF16TAB=1): gather straight from the bf16 scoring table (288 B rows) instead of a widened
            # fp32 copy (512 B rows): the same values and no second 512 MB table in HBM -- but measured SLOWER at config 2
            # (2.58 vs 2.35 ms: the gather is bound by requests in flight, not bytes, and the 288-byte pitch splits rows
            # over one more 128-byte line), so the widened copy stays the default
            cabi.call("htcn_sampled_rank_loss_wt", scores.hout.data_ptr(), Q, self.wt.data_ptr(), scores.y_rows.data_ptr(),
                      neg.data_ptr(), k, cabi.LOSS_KINDS[kind], float(a.hinge_delta), float(a.nce_weight),
                      int(a.num_neg_sample), out.data_ptr(), self.stream_ptr())
            return out
        if self.wt_f32 is None:                  # fp32 gather table
            self.wt_f32 = self.wt[:, :D].float().contiguous()
        cabi.call("htcn_sampled_rank_loss", scores.hout.data_ptr(), self.act_dtype, Q, self.wt_f32.data_ptr(),
                  scores.y_rows.data_ptr(), neg.data_ptr(), k, cabi.LOSS_KINDS[kind], float(a.hinge_delta),
                  float(a.nce_weight), int(a.num_neg_sample), out.data_ptr(), self.stream_ptr())
        return out

    # ------------------------------------------------------------------ the reference's sess.run
    def step_async(self, x_list=None, y_list=None, mask_list=None, state=None, metrics=True, per_position=False, topk=0,
                   state_on_device=False, neg_ids=None, staged=None, mask_warmstart=None, x_gap=None):
        """Enqueue one step (H2D on the copy stream, kernels and the D2H of the results on the compute stream) and
        return a ``PendingStep``; ``.result()`` waits for it and returns the host dict of ``step``.  Submitting step
        i+1 before collecting step i hides the host-side batch packing and the PCIe copies behind the kernels."""
        torch = _torch()
        if staged is None:         # host batch; a DeviceBatcher (hiertcn_b200.device_batcher) hands in `staged` directly
            staged = self.stage(x_list, y_list, mask_list, state, neg_ids, mask_warmstart, x_gap)
        scores, state_out = self.forward(staged=staged)
        if "neg_ids" in staged:
            neg_ids = staged["neg_ids"]
        if topk and metrics:        # prime the cache with the fused loss + rank + top-k sweeps (2 instead of 3)
            fused = self.score(scores, ce=True, rank=True, topk=topk)
            scores._cache[(True, True, 0)] = fused
            scores._cache[(False, False, topk)] = fused
        r = self.loss(scores, metrics=metrics, per_position=per_position)
        dev = {"scalars": r["scalars"]}
        if not state_on_device:
            dev["state"] = state_out
        for n in ("loss_bt", "ranks", "ranks_float"):
            if n in r:
                dev[n] = r[n]
        if neg_ids is not None:                     # sampled ranking loss of reference loss.py:22-71 on the same forward
            dev["sampled"] = self.sampled_loss_mean(scores, neg_ids)
        if topk:
            t = self.score(scores, ce=False, rank=False, topk=topk)
            if t:
                dev["topk_val"], dev["topk_idx"] = t["topk_val"], t["topk_idx"]
            dev["row_of"] = scores.row_of
        slot = 2 + self._res_next                   # result staging sets 2..4 (0/1 are the input sets)
        self._res_next = (self._res_next + 1) % 3
        host = {}
        for k, t in dev.items():
            pt, _ = self._pinned(slot, "res_" + k, tuple(t.shape), t.dtype)
            pt.copy_(t, non_blocking=True)
            host[k] = pt
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        return PendingStep(ev, host, state_out if state_on_device else None, topk, self.layout_meta)

    def step(self, x_list=None, y_list=None, mask_list=None, state=None, metrics=True, per_position=False, topk=0,
             state_on_device=False, neg_ids=None, staged=None, mask_warmstart=None, x_gap=None):
        """Host in, host out -- the call ``sess.run([loss, state, ranks_float, ...], feed_dict)`` of
        run_hier_xing.py:145-149 maps to.  Includes the H2D of the batch and the D2H of the results.
        ``state`` may be a numpy array (reference behaviour) or the device tensor returned by a previous step with
        ``state_on_device=True`` (then the carried state never leaves HBM)."""
        return self.step_async(x_list, y_list, mask_list, state, metrics, per_position, topk, state_on_device,
                               neg_ids, staged, mask_warmstart, x_gap).result()


class PendingStep:
    """Handle of an enqueued step; ``result()`` blocks on its completion event and copies the results out of the
    pinned staging buffers (which are reused three steps later)."""

    def __init__(self, event, host, state_dev, topk, layout_meta=None):
        self.event, self.host, self.state_dev, self.topk, self.layout_meta = event, host, state_dev, topk, layout_meta

    def result(self):
        self.event.synchronize()
        h = {k: v.numpy() for k, v in self.host.items()}
        sc = h["scalars"]
        out = dict(loss=sc[0], recall1=sc[1], recall5=sc[2], recall10=sc[3], mrr=sc[4], mrp=sc[5],
                   user_count=sc[6], n_valid=sc[7])
        if self.state_dev is not None:
            out["state"] = self.state_dev              # device layout [B, G*128] (opaque; feed it back as ``state``)
        else:
            from .weights import unpad_state
            st = h["state"].copy()
            out["state"] = unpad_state(st, self.layout_meta) if self.layout_meta is not None else st
        for n in ("loss_bt", "ranks", "ranks_float", "row_of"):
            if n in h:
                out[n] = h[n].copy()
        if "sampled" in h:
            out["sampled_loss"] = float(h["sampled"][0])
        if self.topk:
            out["topk_val"] = h["topk_val"].copy() if "topk_val" in h else np.zeros((0, self.topk), np.float32)
            out["topk_idx"] = h["topk_idx"].copy() if "topk_idx" in h else np.zeros((0, self.topk), np.int32)
        return out


# ---------------------------------------------------------------------- functional surface
_MODELS = {}


def _weights_fingerprint(weights):
    """cheap content key of a weight dict: names, shapes and a strided sample of every tensor (a full sha256 of a 1M-item
    table costs a second per call); catches a dict mutated in place or replaced by another one at a recycled id()"""
    if weights is None:
        return None
    h = []
    for k in sorted(weights):
        a = np.asarray(weights[k])
        flat = a.reshape(-1)
        step = max(1, flat.size // 61)
        h.append((k, a.shape, flat[::step][:64].astype(np.float64).tobytes()))
    return hash(tuple(h))


def _model_for(args, weights, precision):
    key = (id(weights), _weights_fingerprint(weights), precision, int(args.item_num), tuple(args.tcn_channel),
           int(args.num_layer), int(args.kernel_size))
    hit = _MODELS.get(key)
    if hit is None:
        m = HierTCN(args, weights, precision=precision).build()
        _MODELS.clear()
        _MODELS[key] = (m, weights)          # the strong reference keeps id(weights) from being recycled
        return m
    return hit[0]


def model_hier(args, x, y, mask, state, x_gap=None, x_impression=None, name="hier", reuse=None, training=True,
               weights=None, precision=None, model=None):
    """Signature of reference model_hier.py:21: x, y are S-lists of id arrays [B, L_s] (the reference feeds
    one-hot tensors built from these ids, model.py:59-61), mask an S-list of [B,1], state [B, G*H].
    Returns (pred_all, state): pred_all is a lazy ``CatalogScores`` (call .materialize() for the dense
    [B,T,N] tensor at small N); state is a numpy array.  ``model``: an explicit built ``HierTCN`` to run on (otherwise
    one is built from ``weights`` and cached).  The returned handle is valid until the next forward on that model."""
    if x_impression is not None:
        raise NotImplementedError("has_impression is unreachable in the XING runner (SURVEY A.8 #12)")
    if name != "hier":
        raise NotImplementedError("variable scope other than 'hier'")
    m = model if model is not None else _model_for(args, weights, precision or getattr(args, "precision", "bf16"))
    scores, state_out = m.forward(x, y, mask, state, x_gap=x_gap)
    from .weights import unpad_state
    return scores, unpad_state(state_out.cpu().numpy(), m.layout_meta)
