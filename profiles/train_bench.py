"""Times one training step (forward with saved activations + backward + Adam) on synthetic XING-shaped batches.
usage: python profiles/train_bench.py [B] [N] [steps] [f32|bf16]"""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from hiertcn_b200 import _cabi as cabi                      # noqa: E402
from hiertcn_b200.args import make_args                      # noqa: E402
from hiertcn_b200.data_loader import synthetic_batch         # noqa: E402
from hiertcn_b200.model_hier import HierTCN                  # noqa: E402
from hiertcn_b200.train import HierTCNTrainer                # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
N = int(sys.argv[2]) if len(sys.argv) > 2 else 20778
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
precision = sys.argv[4] if len(sys.argv) > 4 else "f32"
a = make_args(["--item_num", str(N), "--batch_size", str(B)])
tr = HierTCNTrainer(HierTCN(a, None, precision=precision).build())
x, y, m = synthetic_batch(B, 10, 20, N, seed=1, lengths="dense", id_dist="zipf")
state = None
for _ in range(2):
    state = tr.train_step(x, y, m, state, state_on_device=True)["state"]
torch.cuda.synchronize()
staged = tr.m.stage(x, y, m, state)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
t_fb = t_opt = 0.0
l0 = cabi.launch_count
for _ in range(steps):
    ev[0].record()
    r = tr.forward_backward(staged=staged)
    ev[1].record()
    tr.apply_gradients(r["scalars"])
    ev[2].record()
    torch.cuda.synchronize()
    t_fb += ev[0].elapsed_time(ev[1])
    t_opt += ev[1].elapsed_time(ev[2])
launches = (cabi.launch_count - l0) // steps
t0 = time.time()
out = tr.train_step(x, y, m, state, state_on_device=True)
torch.cuda.synchronize()
wall = (time.time() - t0) * 1e3
print(json.dumps(dict(B=B, N=N, precision=precision, fwd_bwd_ms=t_fb / steps, adam_ms=t_opt / steps, e2e_wall_ms=wall, launches=launches,
                      user_seq_per_s=B / ((t_fb + t_opt) / steps / 1e3), loss=out["loss"])))
