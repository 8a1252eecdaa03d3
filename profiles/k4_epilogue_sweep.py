"""K4 CE+rank sweep at cfg2 size under the epilogue variants selected by HTCN_K4_EPI (read per call by libhtcn.so).

    python profiles/k4_epilogue_sweep.py [--variants -1,0,2,4,...] [--out gpurun_out/k4_epi.jsonl]

-1 = scalar epilogue (FFMA/FADD per logit, 1 of 8 exponentials on the FMA pipe); n >= 0 = packed f32x2 epilogue with n of
every 16 logit pairs evaluated by the packed polynomial; +100 = degree-2 polynomial.  Prints, per variant, the median
sweep time, TFLOP/s, the loss (must agree between variants to ~1e-5) and the SM clock sampled during the run."""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hiertcn_b200.args import make_args  # noqa: E402
from hiertcn_b200.data_loader import synthetic_batch  # noqa: E402
from hiertcn_b200.model_hier import HierTCN  # noqa: E402
from hiertcn_b200.weights import hier_weight_shapes, init_weights  # noqa: E402


class Clocks(threading.Thread):
    def __init__(self):
        super().__init__(daemon=True)
        self.samples, self.stop = [], False

    def run(self):
        while not self.stop:
            try:
                o = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-i", "0"],
                                   capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append((float(o[0]), float(o[1])))
            except Exception:
                pass
            time.sleep(0.1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variants", default="-1,0,2,3,4,5,6,104,105,106")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "k4_epi.jsonl"))
    ap.add_argument("--iters", type=int, default=3)
    opt = ap.parse_args()
    os.makedirs(os.path.dirname(opt.out), exist_ok=True)
    out = open(opt.out, "w")
    N, B, S, L = 1_000_000, 4096, 10, 20
    a = make_args(["--item_num", str(N), "--batch_size", str(B)])
    w = init_weights(hier_weight_shapes(N), seed=1234, kernel_scale=2.0, bias_noise=0.1)
    model = HierTCN(a, w, precision="bf16").build()
    x, y, m = synthetic_batch(B, S, L, N, seed=1, lengths="dense", id_dist="uniform")
    d = model.stage(x, y, m, None)
    scores, _ = model.forward(staged=d)
    Q = scores.Q
    flops = 2.0 * Q * 128 * N
    for v in [int(s) for s in opt.variants.split(",")]:
        if v >= 2000:                        # 2000 + n: the folded sweep (the default) with n polynomial pairs per 16
            os.environ.pop("HTCN_K4_EPI", None)
            os.environ["HTCN_K4_FOLD_POLY"] = str(v - 2000)
        else:
            os.environ["HTCN_K4_EPI"] = str(v)
        model.sweep_events = []
        clk = Clocks()
        res = None
        for it in range(opt.iters + 1):
            if it == 1:
                clk.start()
            scores._cache.clear()
            res = model.score(scores, ce=True, rank=True)
            torch.cuda.synchronize()
        clk.stop = True
        clk.join()
        ev = model.sweep_events[-opt.iters:]
        ms = float(np.median([e0.elapsed_time(e1) for e0, e1, _ in ev]))
        model.sweep_events = None
        loss = float(res["loss_row"].double().mean().item())          # mean row loss: must agree between variants
        rsum = int(res["rank_row"].long().sum().item())               # must be identical between variants
        sm = [s for s, _ in clk.samples]
        pw = [p for _, p in clk.samples]
        rec = dict(variant=v, ms=ms, tflops=flops / ms / 1e9, loss=loss, rank_sum=rsum,
                   sm_mhz=float(np.median(sm)) if sm else None, power_w=float(np.median(pw)) if pw else None)
        print(json.dumps(rec), flush=True)
        out.write(json.dumps(rec) + "\n")
        out.flush()


if __name__ == "__main__":
    main()
