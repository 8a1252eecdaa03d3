"""python profiles/k2_compaction_check.py -- K2 at the config-2 shape with and without output compaction (out_row) and sbias,
dense and ragged session lengths: which part of the step's K2 call is not in kernel_bench.py's number?"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from hiertcn_b200 import _cabi as cabi
from hiertcn_b200.args import make_args
from hiertcn_b200.data_loader import synthetic_batch
from hiertcn_b200.model_hier import HierTCN
sys.path.insert(0, "profiles")
from kernel_bench import timed

N, B, S, L = 100_000, 4096, 10, 20
T = S * L
model = HierTCN(make_args(["--item_num", str(N), "--batch_size", str(B)]), None, precision="bf16").build()
st = model.stream_ptr()
for lengths in ("dense", "ragged"):
    x, y, m = synthetic_batch(B, S, L, N, seed=1, lengths=lengths, id_dist="uniform")
    d = model.stage(x, y, m, None)
    slot_p, keep = cabi.int_array(d["slot_off"])
    xe = torch.randn((B * T, 128), device="cuda").to(torch.bfloat16)
    sbias = torch.randn((S, B, 128), device="cuda")
    hout = torch.empty((B * T, 128), dtype=torch.bfloat16, device="cuda")
    sc = torch.empty((cabi.tcn_scratch_floats(2, 5),), dtype=torch.float32, device="cuda")
    for name, sb, ro in (("plain", None, None), ("sbias", sbias, None), ("sbias+out_row", sbias, d["row_of"])):
        ms = timed(lambda: cabi.call("htcn_tcn_forward", xe.data_ptr(), cabi.HTCN_BF16, cabi.HTCN_BF16, model.w_in_x.data_ptr(),
                                     sb.data_ptr() if sb is not None else None, model._conv_w_pp[0], model._conv_b_pp[0], None, None, 2, 5,
                                     slot_p, B, T, S, ro.data_ptr() if ro is not None else None, hout.data_ptr(), cabi.HTCN_BF16,
                                     sc.data_ptr(), st))
        print("%-7s %-14s %.4f ms  (Q = %d of %d positions)" % (lengths, name, ms, d["Q"], B * T))
