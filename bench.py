#!/usr/bin/env python
"""Benchmark of the HierTCN hot path: user-sequences/s for forward + full-catalog scoring.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one pass of the hot path over one batch of synthetic XING-shaped input:
K1 gather+meanpool -> K3 GRU over sessions -> K2 causal conv stack -> sampled ranking loss (k=20 negatives,
hinge_logsigmoid, loss.py:42-50) -> K4 full-catalog scoring with fused softmax-CE + rank metrics -> masked
two-level means (the fetch list of run_hier_xing.py:145-149).
Workload at every N: BASELINE.json configs[1] -- batch 4096 users x 10 sessions x 20 positions (dense),
~1M items, emb 100-d (zero-padded to 128), bf16 tier; users shard data-parallel over ranks (weak scaling),
the catalog is replicated, the only collective is the all-reduce of the loss/metric partial sums.

Prints ONE JSON line (rank 0).  `value` = device-resident throughput; `e2e` = through HierTCN.step with host
numpy buffers (H2D of the batch + D2H of loss/metrics/state inside the timed region).
`--impl reference` times the CPU port of the reference graph (oracle/torch_cpu.py; TensorFlow 1.6 cannot be
installed here) on the host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

# stdout carries exactly ONE JSON line.  NCCL prints "NCCL version ..." to stdout under NCCL_DEBUG=VERSION and honours
# NCCL_DEBUG_FILE only above that level, so VERSION is raised to WARN (same line, now into the file = stderr).
if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1]: the configuration the metric is quoted on
    "cfg2": dict(B=4096, S=10, L=20, N=1_000_000, emb_dim=100, lengths="dense",
                 desc="HierTCN fwd + full-catalog CE+rank scoring, XING shape, 1M items, emb 100-d, batch 4096 users"),
    # BASELINE.json configs[0]: the reference's own CPU-runnable case (parity-test size)
    "cfg1": dict(B=64, S=10, L=20, N=20778, emb_dim=128, lengths="dense",
                 desc="HierTCN fwd + full-catalog CE+rank scoring, XING shape, 20778 items, batch 64 users"),
    # BASELINE.json configs[4]: training step (fwd + bwd + Adam), data-parallel over users with an NCCL gradient
    # all-reduce, GLOBAL batch 4096 (strong scaling), full-softmax CE at the reference's catalog size
    "cfg5": dict(B=4096, S=10, L=20, N=20778, emb_dim=128, lengths="dense", train=True,
                 desc="HierTCN training step (fwd+bwd+Adam, full-softmax CE), XING shape, 20778 items, global batch 4096 users"),
}
METRIC = "user-seqs/sec HierTCN fwd+full-catalog scoring"
UNIT = "user-seq/s"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sust=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback")


def load_traffic(opt, wl):
    """dram bytes per launch of the dominant kernel from the committed ncu capture (only valid for the default
    cfg2 / bf16 configuration it was taken on); None otherwise"""
    p = os.path.join(ROOT, "profiles", "r1_k4_traffic.json")
    if opt.workload != "cfg2" or opt.batch or opt.precision != "bf16" or opt.n_split or not os.path.exists(p):
        return None
    d = json.load(open(p))
    return d["dram_bytes_read"] + d["dram_bytes_write"]


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], 0.0, set()
        for t, line in self.rows:
            f = [v.strip() for v in line.split(",")]
            if len(f) < 8 or not (t0 - 0.05 <= t <= t1 + 0.15):
                continue
            try:
                sm.append(float(f[1]))
                mx = max(mx, float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def make_inputs(wl, seed, B=None):
    from hiertcn_b200.data_loader import ItemSampler, synthetic_batch
    B = B or wl["B"]
    smp = make_inputs.sampler.get(wl["N"])
    if smp is None:
        smp = make_inputs.sampler[wl["N"]] = ItemSampler(wl["N"], "zipf")
    x, y, m = synthetic_batch(B, wl["S"], wl["L"], wl["N"], seed=seed, lengths=wl["lengths"], sampler=smp)
    s0 = np.random.default_rng(seed + 7).normal(0, 0.5, size=(B, 256)).astype(np.float32)
    return x, y, m, s0


make_inputs.sampler = {}


def make_weights(wl):
    from hiertcn_b200.weights import hier_weight_shapes, init_weights
    return init_weights(hier_weight_shapes(wl["N"], emb_dim=wl["emb_dim"]), seed=1234, kernel_scale=2.0)


def cpu_baseline(wl, w, seconds_target=15.0, steps=1, warmup=0):
    """Oracle port on the host cores (all of them), bounded sample of the same workload."""
    import torch
    from oracle.torch_cpu import CpuHierTCN          # the one CPU leg allowed to execute oracle/
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = CpuHierTCN(w)
    # calibrate the sample: ~51 kFLOP/item/user-seq -> pick users so one step is ~seconds_target
    flop_per_user = 51200.0 * wl["N"] + 9.0e7
    users = int(max(1, min(wl["B"], seconds_target * 1.5e11 * min(cores, 32) / 32 / flop_per_user)))
    x, y, m, s0 = make_inputs(wl, seed=99, B=users)
    for _ in range(warmup):
        model.step(x, y, m, s0)
    t0 = time.perf_counter()
    for _ in range(steps):
        out = model.step(x, y, m, s0)
    dt = (time.perf_counter() - t0) / steps
    return dict(value=users / dt, unit=UNIT, cores=cores, kind="port",
                sample="%d of %d users per step, full %d-item catalog, fp32, torch-CPU port of the TF graph "
                       "(gather + streamed catalog), %.1f s/step" % (users, wl["B"], wl["N"], dt)), dt, users, out


def run_reference(opt, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = make_weights(wl)
    base, dt, users, _ = cpu_baseline(wl, w, seconds_target=12.0, steps=opt.steps, warmup=min(opt.warmup, 1))
    line = {"metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": opt.gpus, "steps": opt.steps,
            "warmup": min(opt.warmup, 1), "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": wl["desc"], "sample_users_per_step": users, "items": wl["N"]},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def run_training(opt, wl):
    """BASELINE config 5: one optimisation step = forward with saved activations + backward + all-reduce + Adam
    (hiertcn_b200.train).  Strong scaling: the global batch is split over the ranks."""
    import torch
    import torch.distributed as dist
    from hiertcn_b200 import _cabi as cabi
    from hiertcn_b200.args import make_args
    from hiertcn_b200.model_hier import HierTCN
    from hiertcn_b200.train import HierTCNTrainer
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    warmup = max(opt.warmup, 3)
    peaks = load_peaks()
    Bg = wl["B"]
    B = Bg // world
    w = make_weights(wl)
    a = make_args(["--item_num", str(wl["N"]), "--batch_size", str(B)])
    tr = HierTCNTrainer(HierTCN(a, w, precision=opt.precision).build(), dist=dist if world > 1 else None, world=world)
    x, y, m, s0 = make_inputs(wl, seed=1 + rank, B=B)
    staged = tr.m.stage(x, y, m, s0)
    sampler = ClockSampler(local)
    if rank == 0 and not os.environ.get("HTCN_BENCH_NO_SAMPLER"):
        sampler.start()
        time.sleep(0.5)

    def dev_step():
        r = tr.forward_backward(staged=staged)
        return tr.apply_gradients(r["scalars"])

    for _ in range(warmup):
        sc = dev_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    l0 = cabi.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t_wall0 = time.time()
    e0.record()
    for _ in range(opt.steps):
        sc = dev_step()
    e1.record()
    torch.cuda.synchronize()
    t_wall1 = time.time()
    launches = cabi.launch_count - l0
    if world > 1:
        dist.barrier()
    t_ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(t_ms.item()) / opt.steps
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    loss_dev = float(sc.cpu().numpy()[0])
    # end to end: numpy batch in (H2D inside), loss + carried state out (D2H inside)
    state = s0
    for _ in range(2):
        state = tr.train_step(x, y, m, state)["state"]
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e2e_steps = max(2, min(opt.steps, 5))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        out = tr.train_step(x, y, m, state)
        state = out["state"]
    dt = time.perf_counter() - t0
    t_e = torch.tensor([dt], device="cuda")
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
        dist.destroy_process_group()
    if rank != 0:
        return
    Q = int(staged["Q"])
    flops = 3 * 2.0 * Q * 128 * wl["N"] * world          # catalog products: logits, dHout, dW_out^T (recompute not counted)
    line = {"metric": "user-seqs/sec HierTCN training step (fwd+bwd+Adam)", "value": Bg / (ms_per_step * 1e-3), "unit": UNIT,
            "n_gpus": world, "steps": opt.steps, "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": opt.precision, "data": "synthetic",
            "config": {"workload": wl["desc"], "users_per_gpu": B, "items": wl["N"], "scored_rows_per_gpu": Q,
                       "parallelism": "dp%d over users, one all-reduce of the flat gradient buffer (%d floats)" % (world, tr.n_flat),
                       "l2": "saved activations %.0f MB per step vs 126 MB L2" % (B * 200 * 512 * 5 / 1e6)},
            "loss": loss_dev, "loss_after_e2e_steps": out["loss"],
            "e2e": {"value": Bg / (float(t_e.item()) / e2e_steps), "unit": UNIT, "h2d_bytes_per_step": int(staged["h2d_bytes"]),
                    "d2h_bytes_per_step": 8 * 4 + B * 256 * 4, "steps": e2e_steps,
                    "api": "HierTCNTrainer.train_step(x_list, y_list, mask_list, state): numpy in / loss + state out"},
            "gpu_launches": int(launches),
            "roofline": {"kernel": "whole step, catalog products only", "bound": "tensor", "achieved": flops / (ms_per_step * 1e-3) / 1e12 / world,
                         "peak": peaks["tf_sust"], "unit": "TFLOP/s", "frac": flops / (ms_per_step * 1e-3) / 1e12 / world / peaks["tf_sust"],
                         "peak_source": peaks["src"], "traffic": None},
            "clocks": clocks}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--batch", type=int, default=0, help="override users per GPU (debugging)")
    ap.add_argument("--n-split", type=int, default=0, help="override the catalog split count of the K4 sweep")
    opt = ap.parse_args()
    wl = dict(WORKLOADS[opt.workload])
    if opt.batch:
        wl["B"] = opt.batch
    if opt.impl == "reference":
        return run_reference(opt, wl)
    if wl.get("train"):
        return run_training(opt, wl)

    import torch
    import torch.distributed as dist
    from hiertcn_b200 import _cabi as cabi
    from hiertcn_b200.args import make_args
    from hiertcn_b200.model_hier import HierTCN

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    warmup = max(opt.warmup, 3)
    peaks = load_peaks()

    w = make_weights(wl)
    a = make_args(["--item_num", str(wl["N"]), "--batch_size", str(wl["B"]), "--emb_dim", str(wl["emb_dim"])])
    model = HierTCN(a, w, precision=opt.precision).build()
    model.force_n_split = opt.n_split
    x, y, m, s0 = make_inputs(wl, seed=1 + rank)
    B, T = wl["B"], wl["S"] * wl["L"]

    # global loss/metrics: all-reduce of (per-rank mean * user_count, user_count) into a persistent buffer.  The
    # collective runs on NCCL's own stream (async_op) and nothing on the compute stream depends on it, so it
    # overlaps the next step; the handles are drained before the timed region closes.
    red_buf = torch.zeros(8, dtype=torch.float32, device="cuda")
    pending = []

    def reduce_scalars(sc):
        if world == 1 or os.environ.get("HTCN_BENCH_NO_ALLREDUCE"):
            return sc
        red_buf.copy_(sc)
        red_buf[:6] *= sc[6]
        pending.append(dist.all_reduce(red_buf, async_op=True))
        return red_buf

    def drain():
        while pending:
            pending.pop().wait()
        if world > 1 and not os.environ.get("HTCN_BENCH_NO_ALLREDUCE"):
            out = red_buf.clone()
            out[:6] /= out[6]
            return out
        return None

    # ---------------- device-resident arm (`value`) ----------------
    staged = model.stage(x, y, m, s0)
    neg_host = np.random.default_rng(11 + rank).integers(1, wl["N"], size=(int(staged["Q"]), 20), dtype=np.int32)
    neg_dev = torch.from_numpy(neg_host).cuda()
    torch.cuda.synchronize()
    sampled = {}

    def dev_step():
        scores, state_out = model.forward(staged=staged)
        sampled["scalars"] = model.sampled_loss_mean(scores, neg_dev)
        r = model.loss(scores, metrics=True)
        return reduce_scalars(r["scalars"])

    # the clock sampler starts BEFORE the warm-up: nvidia-smi's start-up (NVML init, ~0.3 s) takes driver locks
    # that stall NCCL launches, so it must not overlap the timed region; only samples inside it are used
    sampler = ClockSampler(local)
    if rank == 0 and not os.environ.get("HTCN_BENCH_NO_SAMPLER"):
        sampler.start()
        time.sleep(0.5)
    for _ in range(warmup):
        sc = dev_step()
    drain()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    model.sweep_events = []
    l0 = cabi.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t_wall0 = time.time()
    e0.record()
    for _ in range(opt.steps):
        sc = dev_step()
    g = drain()
    sc = g if g is not None else sc
    e1.record()
    torch.cuda.synchronize()
    t_wall1 = time.time()
    launches = cabi.launch_count - l0
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    sweep = [(a0.elapsed_time(a1), fl) for a0, a1, fl in model.sweep_events]
    model.sweep_events = None
    t_ms = torch.tensor([ms], device="cuda")
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_per_step = float(t_ms.item()) / opt.steps
    value = world * B / (ms_per_step * 1e-3)
    scalars = sc.cpu().numpy()

    # ---------------- end-to-end arm (`e2e`): host numpy in, host results out ----------------
    for _ in range(2):
        out = model.step(x, y, m, s0, neg_ids=neg_host)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e2e_steps = max(2, min(opt.steps, 5))
    t0 = time.perf_counter()
    pend = None
    for _ in range(e2e_steps):                  # depth-2 software pipeline: pack + H2D of step i+1 overlap step i
        nxt = model.step_async(x, y, m, s0, neg_ids=neg_host)
        if pend is not None:
            out = pend.result()
        pend = nxt
    out = pend.result()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t_e = torch.tensor([dt], device="cuda")
    if world > 1:
        dist.all_reduce(t_e, op=dist.ReduceOp.MAX)
    e2e_value = world * B / (float(t_e.item()) / e2e_steps)
    h2d = staged["h2d_bytes"] + neg_host.nbytes
    d2h = 8 * 4 + B * 256 * 4

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---------------- roofline of the dominant kernel (K4 sweep) ----------------
    sweep_ms = float(np.mean([s[0] for s in sweep]))
    achieved = sweep[0][1] / (sweep_ms * 1e-3) / 1e12
    peak = peaks["tf_sust"]
    roofline = {"kernel": "k4_score_bf16_cg2<CE|RANK, packed f32x2 epilogue>" if opt.precision == "bf16" else "k4_score_f32",
                "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "peak_source": peaks["src"] + " (sustained cuBLAS bf16: the kernel is timed inside a long step)",
                "traffic": load_traffic(opt, wl), "ms_per_launch": sweep_ms, "share_of_step": sweep_ms / ms_per_step,
                "flops_per_launch": sweep[0][1]}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": opt.steps, "warmup": warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": opt.precision, "data": "synthetic",
            "config": {"workload": wl["desc"], "users_per_gpu": B, "sessions": wl["S"], "positions": wl["L"],
                       "items": wl["N"], "emb_dim": wl["emb_dim"], "scored_rows_per_gpu": int(staged["Q"]),
                       "parallelism": "dp%d over users, catalog replicated" % world,
                       "l2": "inputs larger than L2: catalog %.0f MB + activations %.0f MB per step vs 126 MB L2"
                             % (wl["N"] * 256 / 1e6, B * T * 256 * 3 / 1e6)},
            "loss": float(scalars[0]), "mrr": float(scalars[4]), "sampled_loss": float(sampled["scalars"][0].item()),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "steps": e2e_steps, "api": "HierTCN.step_async(x_list, y_list, mask_list, state).result(): numpy in / numpy out, pipelined 2 deep"},
            "gpu_launches": int(launches), "roofline": roofline, "clocks": clocks}
    if world > 1:
        dist.destroy_process_group()
    if not opt.no_cpu_baseline and world == 1:          # the CPU baseline is reported at N=1 only
        base, _, _, _ = cpu_baseline(wl, w, seconds_target=12.0)
        line["cpu_baseline"] = base
    print(json.dumps(line))


if __name__ == "__main__":
    main()
