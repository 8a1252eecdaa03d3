"""torch-CPU port of the hot path for the CPU baseline legs of bench.py (TEST / BASELINE INFRASTRUCTURE --
see oracle/__init__.py; parity unpinned for TF op semantics).

The reference is a TensorFlow-1.6 graph and TensorFlow cannot be installed here, so the "reference arm"
is this port: the same op sequence as oracle/hiertcn_oracle.py (which is checked against the golden
vectors produced by the reference's own python), expressed with multi-threaded torch-CPU kernels so it
uses every host core.  Two forms:

* ``literal=True``  -- one-hot x table matmuls and a materialised [B,T,N] logits tensor, like the TF graph
                       (model.py:59-61, model_hier.py:39-94, loss.py:20-21,163-221).  Small catalogs only.
* ``literal=False`` -- gather + hoisted GRU + catalog streamed in chunks (what a careful CPU implementation
                       would do; needed at N ~ 1M where [B,T,N] does not fit).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def _w(w, k):
    return torch.from_numpy(np.ascontiguousarray(w[k], dtype=np.float32))


class CpuHierTCN:
    def __init__(self, w, num_layer=2, n_levels=2, kernel_size=5):
        self.G, self.n_levels, self.K = num_layer, n_levels, kernel_size
        self.E = _w(w, "hier/emb/kernel")
        self.be = _w(w, "hier/emb/bias")
        self.w_in = _w(w, "hier/tcn/emb/kernel")
        self.D = self.E.shape[1]
        self.conv = [(_w(w, f"hier/tcn/temporal_conv_net/tblock_{l}/conv1/kernel"),
                      _w(w, f"hier/tcn/temporal_conv_net/tblock_{l}/conv1/bias")) for l in range(n_levels)]
        self.gru = [tuple(_w(w, f"hier/multi_rnn_cell/cell_{g}/gru_cell/{n}") for n in
                          ("gates/kernel", "gates/bias", "candidate/kernel", "candidate/bias")) for g in range(num_layer)]
        self.w_out = _w(w, "hier/tcn/dense/kernel")
        self.b_out = _w(w, "hier/tcn/dense/bias")
        self.N = self.w_out.shape[1]

    def _emb(self, ids, literal):
        if literal:                                                          # model.py:59-61 + model_hier.py:50
            oh = F.one_hot(ids, self.N).to(torch.float32) * torch.sign(ids).unsqueeze(-1).to(torch.float32)
            return oh @ self.E
        out = self.E[ids]
        return out * (ids > 0).unsqueeze(-1)

    def _gru(self, x, state):                                                # customed_gru_cell.py:309-337,1050-1073
        H = state.shape[1] // self.G
        new = []
        cur = x
        for g, (wg, bg, wc, bc) in enumerate(self.gru):
            h = state[:, g * H:(g + 1) * H]
            v = torch.sigmoid(torch.cat([cur, h], 1) @ wg + bg)
            r, u = v[:, :H], v[:, H:]
            c = torch.tanh(torch.cat([cur, r * h], 1) @ wc + bc)
            cur = u * h + (1 - u) * c
            new.append(cur)
        return torch.cat(new, 1)

    def _tcn(self, h):                                                       # customized_tcn_cell.py:46-49,109-127
        for lvl, (wk, bk) in enumerate(self.conv):
            d = 2 ** lvl
            xp = F.pad(h.transpose(1, 2), ((self.K - 1) * d, 0))            # causal left pad
            a = F.conv1d(xp, wk.permute(2, 1, 0).contiguous(), bk, dilation=d).transpose(1, 2)
            h = torch.relu(torch.relu(a) + h)
        return h

    @torch.no_grad()
    def step(self, x_list, y_list, mask_list, state, literal=False, chunk=65536):
        """forward + softmax-CE + rank metrics; returns dict(loss, mrr, state)."""
        state = torch.from_numpy(np.ascontiguousarray(state, dtype=np.float32))
        houts = []
        for s in range(len(x_list)):
            x = torch.from_numpy(np.asarray(x_list[s]).astype(np.int64))
            y = torch.from_numpy(np.asarray(y_list[s]).astype(np.int64))
            xe = self._emb(x, literal)
            feat = state.unsqueeze(1).expand(-1, xe.shape[1], -1)
            h0 = torch.cat([xe, feat], -1) @ self.w_in                       # model_hier.py:54-55, model_tcn.py:35
            houts.append(self._tcn(h0))
            if literal:                                                      # model_hier.py:83-85
                oh = F.one_hot(y, self.N).to(torch.float32) * torch.sign(y).unsqueeze(-1).to(torch.float32)
                cnt = torch.sign(oh.abs().sum(2)).sum(1, keepdim=True)
                ys = (oh.sum(1) / cnt) @ self.E + self.be
            else:
                ye = self._emb(y, False)
                ys = ye.sum(1) / (y > 0).sum(1, keepdim=True) + self.be
            state = self._gru(ys, state) * torch.from_numpy(np.asarray(mask_list[s], dtype=np.float32))
        hout = torch.cat(houts, 1)                                           # [B,T,C]
        y_id = torch.from_numpy(np.concatenate([np.asarray(v) for v in y_list], 1).astype(np.int64))
        B, T = y_id.shape
        mask = (y_id > 0).to(torch.float32)
        hq = hout.reshape(B * T, -1) * mask.reshape(-1, 1)                   # pred *= mask_y (model.py:105)
        bq = mask.reshape(-1, 1)                                             # masked rows see zero logits
        zy = ((hq * self.w_out.t()[y_id.reshape(-1)]).sum(1) + self.b_out[y_id.reshape(-1)] * mask.reshape(-1))
        yq = y_id.reshape(-1)
        if literal:
            z = (hq @ self.w_out + self.b_out) * bq
            zy = z.gather(1, yq.unsqueeze(1)).squeeze(1)                     # reduce_sum(score*y) (loss.py:179)
            lse = torch.logsumexp(z, 1)
            rank = (z > zy.unsqueeze(1)).sum(1).to(torch.float32)
        else:                                                                # stream the catalog
            m = torch.full((B * T,), -float("inf"))
            ssum = torch.zeros(B * T)
            rank = torch.zeros(B * T)
            for j0 in range(0, self.N, chunk):
                z = (hq @ self.w_out[:, j0:j0 + chunk] + self.b_out[j0:j0 + chunk]) * bq
                mc = torch.maximum(m, z.max(1).values)
                ssum = ssum * torch.exp(m - mc) + torch.exp(z - mc.unsqueeze(1)).sum(1)
                m = mc
                rank += (z > zy.unsqueeze(1)).sum(1)
                own = (yq >= j0) & (yq < j0 + z.shape[1])                    # a separately computed target logit can
                rows = own.nonzero().squeeze(1)                              # differ from the swept one in the last
                rank[rows] -= (z[rows, yq[rows] - j0] > zy[rows]).to(torch.float32)   # bit: never count the target itself
            lse = m + torch.log(ssum)
        loss_bt = ((lse - zy) * mask.reshape(-1)).reshape(B, T)
        act = mask.sum(1)
        uc = torch.sign(act).sum()
        act = act + 1e-6
        loss = (loss_bt.sum(1) / act).sum() / uc
        rr = ((1.0 / (1.0 + rank)) * mask.reshape(-1)).reshape(B, T)
        mrr = (rr.sum(1) / act).sum() / uc
        return dict(loss=float(loss), mrr=float(mrr), state=state.numpy(), ranks=(rank * mask.reshape(-1)).reshape(B, T).numpy(),
                    loss_bt=loss_bt.numpy())
