#!/usr/bin/env python
"""Per-kernel roofline measurements (device-resident, CUDA events, L2 flushed between iterations).

    python profiles/kernel_bench.py [--out profiles/r2_kernels.jsonl]

One JSON line per kernel with achieved GB/s or TFLOP/s against MEASURED_PEAKS.json:
  K1 gather+meanpool   cfg2 shape, uniform ids over a 1M x 128 fp32 table (512 MB >> L2): HBM-bound
  K3 GRU over sessions cfg2 shape
  K2 conv stack        cfg2 (hier, 2 levels, ragged-free) and cfg3 (TCN only, 4096 x 256, 4 levels), bf16 tcgen05
  K4 sweep             cfg2 CE+rank, rank only, and top-100 (cfg4 per-GPU shape: Q = 4096 queries x 1M-item shard)
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hiertcn_b200 import _cabi as cabi  # noqa: E402
from hiertcn_b200.args import make_args  # noqa: E402
from hiertcn_b200.data_loader import synthetic_batch  # noqa: E402
from hiertcn_b200.model_hier import HierTCN  # noqa: E402

PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) \
    else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}
flush_buf = None


def timed(fn, iters=5, warm=2):
    global flush_buf
    if flush_buf is None:
        flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")     # > 126 MB L2
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush_buf.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def line(out, kernel, ms, bound, work, peak_key, note):
    unit = "GB/s" if bound == "hbm" else "TFLOP/s"
    achieved = work / (ms * 1e-3) / (1e9 if bound == "hbm" else 1e12)
    peak = PEAKS[peak_key]
    rec = dict(kernel=kernel, ms=ms, bound=bound, achieved=achieved, peak=peak, unit=unit, frac=achieved / peak,
               peak_source=peak_key + " (MEASURED_PEAKS.json)", work=work, note=note)
    print(json.dumps(rec))
    out.write(json.dumps(rec) + "\n")
    out.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r2_kernels.jsonl"))
    ap.add_argument("--only", default="")
    opt = ap.parse_args()
    out = open(opt.out, "w")
    N, B, S, L = 1_000_000, 4096, 10, 20
    T = S * L
    a = make_args(["--item_num", str(N), "--batch_size", str(B)])
    from hiertcn_b200.weights import hier_weight_shapes, init_weights
    w = init_weights(hier_weight_shapes(N), seed=1234, kernel_scale=2.0, bias_noise=0.1)   # biases != 0: no all-tie rows
    model = HierTCN(a, w, precision="bf16").build()
    x, y, m = synthetic_batch(B, S, L, N, seed=1, lengths="dense", id_dist="uniform")
    d = model.stage(x, y, m, None)
    st = model.stream_ptr()
    f32, bf16 = torch.float32, torch.bfloat16
    slot_p, keep = cabi.int_array(d["slot_off"])
    want = lambda k: not opt.only or k in opt.only  # noqa: E731

    # ---- K1
    xe = torch.empty((B * T, 128), dtype=bf16, device="cuda")
    xe32 = torch.empty((B * T, 128), dtype=f32, device="cuda")
    yp = torch.empty((S, B, 128), dtype=f32, device="cuda")
    if want("k1"):
        for name, buf, dt, e in (("K1 gather+meanpool (fp32 out)", xe32, cabi.HTCN_F32, 4), ("K1 gather+meanpool (bf16 out)", xe, cabi.HTCN_BF16, 2)):
            ms = timed(lambda: cabi.call("htcn_gather_meanpool", model.E.data_ptr(), model.emb_pitch, model.b_emb.data_ptr(), N, d["x_id"].data_ptr(),
                                         d["y_id"].data_ptr(), slot_p, B, T, S, buf.data_ptr(), dt, yp.data_ptr(), st))
            work = B * (2 * T * 512 + 2 * T * 4 + T * 128 * e + S * 512)       # rows read (x,y) + ids + outputs
            line(out, name, ms, "hbm", work, "hbm_gbs", "cfg2 shape, uniform ids, 1M x 128 fp32 table; two launches (gather, meanpool)")
    cabi.call("htcn_gather_meanpool", model.E.data_ptr(), model.emb_pitch, model.b_emb.data_ptr(), N, d["x_id"].data_ptr(), d["y_id"].data_ptr(),
              slot_p, B, T, S, xe.data_ptr(), cabi.HTCN_BF16, yp.data_ptr(), st)

    # ---- K3
    sbias = torch.empty((S, B, 128), dtype=f32, device="cuda")
    state_out = torch.empty((B, 256), dtype=f32, device="cuda")
    g = model._gru_pp
    k3s = torch.empty(cabi.gru_scratch_bytes(B) // 4, dtype=f32, device="cuda")
    k3 = lambda prec: cabi.call("htcn_gru_sessions", yp.data_ptr(), d["mask"].data_ptr(), d["state"].data_ptr(), g[0][0], g[1][0],  # noqa: E731
                                g[2][0], g[3][0], 2, model.w_in_state.data_ptr(), B, S, prec, k3s.data_ptr(), None, sbias.data_ptr(),
                                state_out.data_ptr(), st)
    if want("k3"):
        for nm, prec in (("fp32 FFMA", cabi.HTCN_F32), ("bf16 tcgen05", cabi.HTCN_BF16)):
            ms = timed(lambda: k3(prec))
            line(out, "K3 GRU over sessions (%s)" % nm, ms, "hbm", B * S * (128 + 2 * 256) * 4, "hbm_gbs",
                 "cfg2 shape; algorithmic bytes only (Yp in, state/sbias out); in practice latency/L2-bound: %.2f TFLOP/s"
                 % (B * (3.93e6 + 10 * 2 * 256 * 128) / (ms * 1e-3) / 1e12))
    k3(cabi.HTCN_BF16)

    # ---- K2
    if want("k2"):
        hout = torch.empty((B * T, 128), dtype=bf16, device="cuda")
        sc = torch.empty((cabi.tcn_scratch_floats(2, 5),), dtype=f32, device="cuda")
        ms = timed(lambda: cabi.call("htcn_tcn_forward", xe.data_ptr(), cabi.HTCN_BF16, cabi.HTCN_BF16, model.w_in_x.data_ptr(),
                                     sbias.data_ptr(), model._conv_w_pp[0], model._conv_b_pp[0], None, None, 2, 5, slot_p, B, T, S, None,
                                     hout.data_ptr(), cabi.HTCN_BF16, sc.data_ptr(), st))
        line(out, "K2 conv stack bf16 tcgen05 (hier, 2 levels)", ms, "tensor", B * (65.54e6 + 2 * T * 128 * 128), "bf16_tflops", "cfg2 shape: 40960 sequences x 20")
        # cfg3: 4096 sequences x 256, 4 levels, dilations 1-2-4-8
        a3 = make_args(["--item_num", "1000", "--tcn_channel", "128,128,128,128"])
        m3 = HierTCN(a3, None, precision="bf16").build()
        B3, L3 = 4096, 256
        xe3 = torch.randn((B3 * L3, 128), device="cuda").to(bf16)
        h3 = torch.empty((B3 * L3, 128), dtype=bf16, device="cuda")
        sc3 = torch.empty((cabi.tcn_scratch_floats(4, 5),), dtype=f32, device="cuda")
        sp3, k3keep = cabi.int_array([0, L3])
        ms = timed(lambda: cabi.call("htcn_tcn_forward", xe3.data_ptr(), cabi.HTCN_BF16, cabi.HTCN_BF16, m3.w_in_x.data_ptr(), None,
                                     m3._conv_w_pp[0], m3._conv_b_pp[0], None, None, 4, 5, sp3, B3, L3, 1, None, h3.data_ptr(), cabi.HTCN_BF16,
                                     sc3.data_ptr(), st))
        line(out, "K2 conv stack bf16 tcgen05 (cfg3: 4096 x 256, 4 levels)", ms, "tensor", B3 * (167.8e6 + 2 * L3 * 128 * 128), "bf16_tflops",
             "BASELINE config 3; useful FLOPs only (the 60-row receptive-field halo recomputed per 128-row tile is not counted)")
        del m3, xe3, h3

    # ---- K4
    if want("k4"):
        scores, _ = model.forward(staged=d)
        Q = scores.Q
        for name, kw, peak in (("K4 sweep CE+rank (cfg2)", dict(ce=True, rank=True), "bf16_tflops_sustained"),
                               ("K4 sweep rank only (cfg2)", dict(ce=False, rank=True), "bf16_tflops_sustained")):
            def run():
                scores._cache.clear()
                model.score(scores, **kw)
            model.sweep_events = []
            ms_total = timed(run, iters=3, warm=1)
            ev = model.sweep_events[-3:]
            ms = float(np.median([e0.elapsed_time(e1) for e0, e1, _ in ev]))
            model.sweep_events = None
            line(out, name, ms, "tensor", 2.0 * Q * 128 * N, peak, "Q = %d rows x %d items; sweep kernel only" % (Q, N))
        # top-100: cfg4 per-GPU shape (4096 queries against a 1M-item shard)
        Qk = 4096
        hq = scores.hout[:Qk].contiguous()
        ns = 9
        tv = torch.empty((ns, Qk, 100), dtype=f32, device="cuda")
        ti = torch.empty((ns, Qk, 100), dtype=torch.int32, device="cuda")
        ov = torch.empty((Qk, 100), dtype=f32, device="cuda")
        oi = torch.empty((Qk, 100), dtype=torch.int32, device="cuda")
        ms = timed(lambda: cabi.call("htcn_score_ce_rank_topk", hq.data_ptr(), cabi.HTCN_BF16, Qk, model.wt.data_ptr(), None, N, 0, None, None, 1,
                                     cabi.SCORE_TOPK, 100, ns, None, None, None, tv.data_ptr(), ti.data_ptr(), st))
        line(out, "K4 heap top-100 sweep (cfg4 per-GPU shape)", ms, "tensor", 2.0 * Qk * 128 * N, "bf16_tflops", "4096 queries x 1M-item shard, 9 splits, one heap per row (robust fallback path)")
        nb = int(cabi.load().htcn_topk_workspace_bytes(cabi.HTCN_BF16, Qk, N, 100, ns))
        ws = torch.empty(nb, dtype=torch.uint8, device="cuda")
        ovf = torch.zeros(1, dtype=torch.int32, device="cuda")
        ms = timed(lambda: cabi.call("htcn_score_topk", hq.data_ptr(), cabi.HTCN_BF16, Qk, model.wt.data_ptr(), None, N, 0, 100, ns, ws.data_ptr(), nb,
                                     ov.data_ptr(), oi.data_ptr(), ovf.data_ptr(), st))
        line(out, "K4 two-pass top-100 (cfg4 per-GPU shape)", ms, "tensor", 2.0 * Qk * 128 * N, "bf16_tflops",
             "4096 queries x 1M-item shard: group-max sweep + threshold select + filter sweep + exact select; useful FLOPs = one GEMM (two are executed); overflow rows %d" % int(ovf.item()))
        ms = timed(lambda: cabi.call("htcn_topk_merge", tv.data_ptr(), ti.data_ptr(), ns, Qk, 100, ov.data_ptr(), oi.data_ptr(), st))
        line(out, "top-k merge (9 parts x 100)", ms, "hbm", ns * Qk * 100 * 8 + Qk * 100 * 8, "hbm_gbs", "k-way merge, rank by counting")


if __name__ == "__main__":
    main()
