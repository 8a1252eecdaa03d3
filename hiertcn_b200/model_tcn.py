"""Single-level TCN (reference model_tcn.py:13-44 as called from model.py:71-73 with the one-hot item sequence;
BASELINE config 3: low-level TCN only, long sequences, dilations 1-2-4-8).

    x ids [B,L] -> 'tcn/emb' (one-hot x [N,128] kernel == a row gather, no bias) -> TemporalConvNet -> dense to N logits

on the same kernels as the hierarchical path: K1 gathers the rows of ``tcn/emb/kernel`` (id 0 -> zeros), K2 runs the
conv stack (its in-projection is fed the identity, because here the 'emb' dense IS the gather), K4 scores the
catalog.  Levels narrower than 128 channels run zero-padded, width changes get the 1x1 down-sample residual of
customized_tcn_cell.py:102-106; levels of 129..256 channels -- the reference's default single-level stack is
[128,128,128,128,256,256] (args.py:310-311) -- run as two 128-wide planes on the fp32 kernels (precision='f32' only; the
tcgen05 tier is built for levels up to 128 channels, config 3 uses [128]*4).
"""
from __future__ import annotations

import numpy as np

from . import _cabi as cabi
from .model_hier import CatalogScores, HierTCN, _torch
from .weights import fold_weightnorm

D = 128


class TCN(HierTCN):
    """model_tcn on one B200: ``forward(x_ids [B,L], y_ids [B,L]) -> CatalogScores`` (losses / ranks / top-k through the
    inherited ``loss`` / ``score`` / ``topk``)."""

    def __init__(self, args, weights, device=None, precision=None, scope="tcn"):
        self.args = args
        self.precision = precision or getattr(args, "precision", "bf16")
        if max(args.tcn_channel) > 256:
            raise NotImplementedError("tcn_channel above 256")
        if max(args.tcn_channel) > 128 and self.precision != "f32":
            raise NotImplementedError("tcn_channel above 128 (the reference's single-level default ends with two 256-channel "
                                      "levels, args.py:310-311) needs precision='f32': two-plane fp32 kernels")
        self.N = int(args.item_num)
        self.K = int(args.kernel_size)
        self.n_levels = len(args.tcn_channel)
        self.G = 0
        self.l2_normalize = bool(getattr(args, "l2_normalize", False))      # model_tcn.py:42-43
        self.host_weights, self.scope, self.device = weights, scope, device
        self.built = False
        self._init_runtime_state()

    def build(self):
        torch = _torch()
        cabi.load()
        if self.device is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        if not cabi.load().htcn_device_ok():
            raise cabi.HtcnError("current CUDA device is not compute capability 10.x (B200)")
        from .weights import to_device_layout
        lay, meta = to_device_layout(fold_weightnorm(self.host_weights), self.scope)
        self.layout_meta = meta
        up = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(self.device)  # noqa: E731
        self.E = up(lay["E"])                                   # [N,128]: the in-projection applied to a one-hot
        self.emb_pitch = D
        self.b_emb, self.w_in_x = up(lay["b_emb"]), up(lay["w_in_x"])
        self.conv_w = [up(lay[f"conv_w{l}"]) for l in range(self.n_levels)]
        self.conv_b = [up(lay[f"conv_b{l}"]) for l in range(self.n_levels)]
        self.ds_w = [up(lay[f"ds_w{l}"]) if meta["ds"][l] else None for l in range(self.n_levels)]
        self.ds_b = [up(lay[f"ds_b{l}"]) if meta["ds"][l] else None for l in range(self.n_levels)]
        self.b_out = up(lay["b_out"])
        w_out = up(lay["w_out"])
        self._finish_build_head(w_out, meta)
        self.wt_f32 = self.wt if self.precision == "f32" else None
        torch.cuda.synchronize(self.device)
        self.refresh_pointer_tables()
        self.built = True
        return self

    def forward(self, x_ids, y_ids=None):
        """x_ids [B,L] item ids (0 = padding), y_ids [B,L] next-item targets (0 = not scored; default: every position
        is scored against id 0, i.e. only hidden states / top-k are meaningful)."""
        if not self.built:
            self.build()
        torch = _torch()
        self.generation += 1
        x = np.ascontiguousarray(np.asarray(x_ids), dtype=np.int32)
        B, L = x.shape
        y = np.ascontiguousarray(np.asarray(y_ids), dtype=np.int32) if y_ids is not None else np.ones_like(x)
        valid = y.reshape(-1) > 0
        row_of = np.where(valid, np.cumsum(valid) - 1, -1).astype(np.int32)
        Q = int(valid.sum())
        dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(self.device)  # noqa: E731
        x_d, y_d, row_d, yrows_d = dev(x), dev(y), dev(row_of), dev(y.reshape(-1)[valid])
        st = self.stream_ptr()
        slot_p, keep = cabi.int_array([0, L])
        xe = self._buf("xe", (B * L, D), self.act_torch_dtype)
        cabi.call("htcn_gather_meanpool", self.E.data_ptr(), D, None, self.N, x_d.data_ptr(), None, slot_p, B, L, 1,
                  xe.data_ptr(), self.act_dtype, None, st)
        hout = self._buf("hout", (self.head_planes * max(Q, 1), D), self.act_torch_dtype)
        prec = self._k2_precision()
        if self.wide:
            scratch = self._buf("k2_scratch", (6 * B * L, D), torch.float32)
            planes_p, planes_keep = cabi.int_array(self.level_planes)
            cabi.call("htcn_tcn_forward_wide", xe.data_ptr(), self.act_dtype, self.w_in_x.data_ptr(), None, self._conv_w_pp[0],
                      self._conv_b_pp[0], self._ds_w_pp[0], self._ds_b_pp[0], planes_p, self.n_levels, self.K, slot_p, B, L, 1,
                      row_d.data_ptr(), hout.data_ptr(), max(Q, 1), scratch.data_ptr(), st)
            return CatalogScores(self, hout, Q, row_d, yrows_d, y_d, B, L)
        if prec == cabi.HTCN_F32:
            scratch = self._buf("k2_scratch", ((3 if self.has_ds else 2) * B * L, D), torch.float32)
        else:
            scratch = self._buf("k2_scratch_bf16", (cabi.tcn_scratch_floats(self.n_levels, self.K),), torch.float32)
        cabi.call("htcn_tcn_forward", xe.data_ptr(), self.act_dtype, prec, self.w_in_x.data_ptr(), None,
                  self._conv_w_pp[0], self._conv_b_pp[0], self._ds_w_pp[0], self._ds_b_pp[0], self.n_levels, self.K, slot_p, B, L, 1, row_d.data_ptr(),
                  hout.data_ptr(), self.act_dtype, scratch.data_ptr(), st)
        return CatalogScores(self, hout, Q, row_d, yrows_d, y_d, B, L)


def model_tcn(args, x, x_gap=None, x_impression=None, name="tcn", reuse=None, training=True, mask=None, weights=None,
              precision=None, y=None):
    """Signature of reference model_tcn.py:13.  ``x`` are item ids [B,L] (the reference is fed their one-hot,
    model.py:56,73).  Returns the lazy ``CatalogScores`` standing for ``pred [B,L,N]``."""
    if x_gap is not None or x_impression is not None:
        raise NotImplementedError("has_gap / has_impression are outside the hot path")
    return TCN(args, weights, precision=precision or getattr(args, "precision", "bf16"), scope=name).build().forward(x, y)
