"""python profiles/k2_stress.py -- randomised shapes: the four-chain conv kernel (all group / stage variants) against the kernels of k2_tcn_bf16.cu, bit for bit."""
import os, sys
import numpy as np, torch
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import test_gpu_kernels as T
from helpers import small_case
from hiertcn_b200 import _cabi as lib
lib.load()
rng = np.random.default_rng(2024)
bad = 0
for case in range(60):
    K = int(rng.integers(1, 6))
    max_lv = 0
    while max_lv < 4 and (K - 1) * (1 << max_lv) <= 32:
        max_lv += 1
    levels = int(rng.integers(1, max_lv + 1)) if K > 1 else int(rng.integers(1, 4))
    S = int(rng.integers(1, 12))
    L = int(rng.choice([1, 2, 5, 20, 33, 100, 127, 128, 129, 200, 300]))
    B = int(rng.choice([1, 2, 3, 7, 40, 150, 330])) if L < 100 else int(rng.choice([1, 2, 5, 40, 160]))
    if L >= 100:
        S = min(S, 2)
    lengths = "ragged" if rng.random() < 0.5 else "dense"
    x, y, m, s0, w = small_case(B=B, S=S, L=L, N=301, seed=case, tcn_channel=(128,) * levels, kernel_size=K, lengths=lengths, kernel_scale=1.0)
    pk = T.pack(x, y, m)
    TT = pk["x_id"].shape[1]
    xe = T.dev(T.O.emb_gather(pk["x_id"], w["hier/emb/kernel"])).to(torch.bfloat16)
    sbias = torch.randn((S, B, 128), device="cuda") * 0.3 if rng.random() < 0.8 else None
    valid = pk["y_id"].reshape(-1) > 0
    row_of = np.where(valid, np.cumsum(valid) - 1, -1).astype(np.int32)
    ro = T.dev(row_of) if rng.random() < 0.5 else None
    n_out = int(valid.sum()) if ro is not None else None
    if n_out == 0:
        ro, n_out = None, None
    res = {}
    for name, env in (("old", {"HTCN_K2_QUAD": "0"}), ("default", {}), ("q1", {"HTCN_K2_QUAD": "1"}), ("q2", {"HTCN_K2_QUAD": "2", "HTCN_K2_FULLTAP": "0"}),
                      ("q2ft", {"HTCN_K2_QUAD": "2", "HTCN_K2_FULLTAP": "1"}), ("q4lag0", {"HTCN_K2_QUAD": "4", "HTCN_K2_LAG": "0"}),
                      ("q4lag3", {"HTCN_K2_QUAD": "4", "HTCN_K2_LAG": "3"}), ("q4i2", {"HTCN_K2_QUAD": "4", "HTCN_K2_ISSUERS": "2"})):
        for k, v in env.items():
            os.environ[k] = v
        try:
            res[name] = T.run_k2_bf16(lib, xe, w, sbias, pk["slot_off"], B, TT, S, K, levels, out_row=ro, n_out=n_out).float().cpu().numpy()
        finally:
            for k in env:
                del os.environ[k]
    for name in res:
        if not np.array_equal(res[name], res["old"]):
            bad += 1
            print("MISMATCH", case, name, dict(B=B, S=S, L=L, K=K, levels=levels, lengths=lengths, compact=ro is not None),
                  np.abs(res[name] - res["old"]).max())
    assert np.isfinite(res["default"]).all()
print("cases 60, mismatches", bad)
