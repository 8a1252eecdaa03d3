#!/usr/bin/env python
"""K3 (GRU over sessions, bf16 tier) -- every kernel variant at the config-2 shape (B = 4096 users, S = 10 sessions):
whole-call time (CUDA events, L2 flushed), agreement with the streaming kernel, one JSON line per variant.

    python profiles/k3_variants.py [--out profiles/r2_k3_variants.jsonl] [--modes 0,1,3,4,5]

HTCN_K3_CLUSTER: 0 = users on M, weights streamed (k3_gru_bf16.cu); 1 / 2 = 4-CTA cluster, resident weight slices, DSMEM
exchange (k3_gru_cluster.cu); 3 (default) / 4 = users on N, no exchange (k3_gru_t.cu: 128 KB of weights resident + 3-stage
ring, 64 KB resident + 7-stage ring); 6 / 7 = its wavefront form (k3_gru_w.cu)."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hiertcn_b200 import _cabi as cabi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r2_k3_variants.jsonl"))
    ap.add_argument("--modes", default="0,1,2,3,4,6,7")
    ap.add_argument("--B", type=int, default=4096)
    ap.add_argument("--S", type=int, default=10)
    ap.add_argument("--iters", type=int, default=20)
    opt = ap.parse_args()
    cabi.load()
    B, S = opt.B, opt.S
    g = torch.Generator(device="cuda").manual_seed(1)
    f32 = torch.float32
    rnd = lambda *s, sc=1.0: (torch.randn(*s, device="cuda", generator=g) * sc).contiguous()  # noqa: E731
    yp, st_in = rnd(S, B, 128, sc=0.5), rnd(B, 256, sc=0.5)
    mask = (torch.rand(S, B, device="cuda", generator=g) > 0.2).to(f32)
    gw = [rnd(256, 256, sc=0.06) for _ in range(2)]
    gb = [rnd(256, sc=0.1) + 1.0 for _ in range(2)]
    cw = [rnd(256, 128, sc=0.06) for _ in range(2)]
    cb = [rnd(128, sc=0.1) for _ in range(2)]
    wis = rnd(256, 128, sc=0.06)
    pps = [cabi.ptr_array([t.data_ptr() for t in ts]) for ts in (gw, gb, cw, cb)]
    scratch = torch.empty(cabi.gru_scratch_bytes(B) // 4, dtype=f32, device="cuda")
    spre = torch.empty((S, B, 256), dtype=f32, device="cuda")
    sbias = torch.empty((S, B, 128), dtype=f32, device="cuda")
    sout = torch.empty((B, 256), dtype=f32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def call(with_pre=True):
        cabi.call("htcn_gru_sessions", yp.data_ptr(), mask.data_ptr(), st_in.data_ptr(), pps[0][0], pps[1][0], pps[2][0], pps[3][0],
                  2, wis.data_ptr(), B, S, cabi.HTCN_BF16, scratch.data_ptr(), spre.data_ptr() if with_pre else None,
                  sbias.data_ptr(), sout.data_ptr(), st)

    ref = None
    out = open(opt.out, "a")
    for mode in opt.modes.split(","):
        os.environ["HTCN_K3_CLUSTER"] = mode
        for t in (spre, sbias, sout):
            t.fill_(7.0)
        call()
        torch.cuda.synchronize()
        res = [t.clone() for t in (spre, sbias, sout)]
        if ref is None:
            ref = res
        diff = max(float((a - b).abs().max()) for a, b in zip(res, ref))
        ts = []
        for _ in range(opt.iters):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            call(with_pre=False)                                   # the inference path does not emit state_pre
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts))
        work = B * S * (128 + 2 * 128) * 4 + 2 * B * 256 * 4       # Yp in, sbias out, state in / out
        rec = dict(kernel="K3 GRU over sessions, HTCN_K3_CLUSTER=%s" % mode, B=B, S=S, ms_per_call=ms, min_ms=float(min(ts)),
                   max_abs_diff_vs_first=diff, algorithmic_gbs=work / (ms * 1e-3) / 1e9,
                   tflops=B * S * 2 * (2 * (256 * 256 + 256 * 128) + 256 * 128) / (ms * 1e-3) / 1e12)
        print(json.dumps(rec))
        out.write(json.dumps(rec) + "\n")
    out.close()


if __name__ == "__main__":
    main()
