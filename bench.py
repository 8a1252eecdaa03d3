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

Extra legs on the same JSON line (every N; the headline `value` stays configs[1], weak scaling):
  `strong`  -- configs[1] with a GLOBAL batch of 4096 users split over the ranks (strong scaling of the forward path)
  `sharded` -- BASELINE configs[3]: CE + rank + top-100 of 4096 queries over 8M items, the catalog sharded N ways
               (hiertcn_b200.dist.ShardedCatalogScorer), device time per phase / collective, and `sharded_parity`:
               ranks, top-k indices and values of the sharded path == the replicated single-GPU path, bit for bit, on a
               1M-item catalog (loss rows to 2e-6)
  `train`   -- BASELINE configs[4]: training step (fwd + bwd + all-reduce + Adam), global batch 4096 (strong scaling),
               with the device time of the gradient all-reduce
  `kernels` -- per-kernel device times of the step (K1 gather, K3 GRU, K2 conv stack, sampled loss) with their rooflines
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
    p = os.path.join(ROOT, "profiles", "r2_k4_traffic.json")
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


def cpu_literal_cfg1(steps=2):
    """BASELINE.md's LITERAL CPU form at configs[0] (B=64, N=20 778): one-hot x table matmuls and a materialised
    [B,T,N] logits tensor, the op sequence of the TF graph (oracle/torch_cpu.py, literal=True), on all host cores."""
    import torch
    from oracle.torch_cpu import CpuHierTCN          # CPU baseline leg only
    wl = WORKLOADS["cfg1"]
    torch.set_num_threads(os.cpu_count() or 1)
    model = CpuHierTCN(make_weights(wl))
    x, y, m, s0 = make_inputs(wl, seed=98)
    model.step(x, y, m, s0, literal=True)
    t0 = time.perf_counter()
    for _ in range(steps):
        model.step(x, y, m, s0, literal=True)
    dt = (time.perf_counter() - t0) / steps
    return {"value": wl["B"] / dt, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
            "sample": "configs[0] whole: %d users, %d items, literal one-hot/[B,T,N] form, %.2f s/step" % (wl["B"], wl["N"], dt)}


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
    """`--workload cfg5`: the training step as the headline line"""
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    line = train_leg(opt, wl, rank, world, local, sample_clocks=True)
    if world > 1:
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line))


def train_leg(opt, wl, rank, world, local, sample_clocks=False):
    """BASELINE config 5: one optimisation step = forward with saved activations + backward + all-reduce + Adam
    (hiertcn_b200.train).  Strong scaling: the global batch is split over the ranks.  Returns the JSON object (rank 0)."""
    import torch
    import torch.distributed as dist
    from hiertcn_b200 import _cabi as cabi
    from hiertcn_b200.args import make_args
    from hiertcn_b200.model_hier import HierTCN
    from hiertcn_b200.train import HierTCNTrainer
    warmup = max(opt.warmup, 3)
    peaks = load_peaks()
    Bg = wl["B"]
    B = Bg // world
    w = make_weights(wl)
    a = make_args(["--item_num", str(wl["N"]), "--batch_size", str(B)])
    tr = HierTCNTrainer(HierTCN(a, w, precision=opt.precision).build(), dist=dist if world > 1 else None, world=world)
    # every rank generates the same global batch and takes its slice of the users
    xg, yg, mg, sg = make_inputs(wl, seed=1, B=Bg)
    lo, hi = rank * B, (rank + 1) * B
    x, y, m, s0 = [v[lo:hi] for v in xg], [v[lo:hi] for v in yg], [v[lo:hi] for v in mg], sg[lo:hi]
    staged = tr.m.stage(x, y, m, s0)
    # in-run parity of the data-parallel exchange: the all-reduced gradient of the rank slices == the gradient one
    # process computes on the whole batch (rank 0 runs it), and every rank holds the same parameters afterwards
    dp_parity = None
    if world > 1:
        r = tr.forward_backward(staged=staged)
        g_dp = tr.grads.clone()
        sc_dp = r["scalars"].clone()
        sc_dp[:6] *= sc_dp[6]
        tr.grads.zero_()
        dist.all_reduce(g_dp)
        dist.all_reduce(sc_dp)
        if rank == 0:
            ref = HierTCNTrainer(HierTCN(a, w, precision=opt.precision).build())
            rr = ref.forward_backward(xg, yg, mg, sg)
            err = float((g_dp - ref.grads).norm() / ref.grads.norm())
            loss_dp, loss_ref = float(sc_dp[0] / sc_dp[6]), float(rr["scalars"][0])
            dp_parity = {"grad_rel_err": err, "loss_dp": loss_dp, "loss_single": loss_ref,
                         # the per-rank sweeps use a different catalog split count than the whole-batch one, so the row
                         # log-sum-exps differ in the last bit and a few bf16-rounded dL/dZ entries land on the neighbouring
                         # bf16 value: ~1e-4 in norm in the bf16 tier (1e-6 in the fp32 tier)
                         "ok": bool(err <= 5e-4 and abs(loss_dp - loss_ref) <= 1e-5 * abs(loss_ref)
                                    and float(sc_dp[6]) == float(rr["scalars"][6]))}
            del ref, rr
        del g_dp
        torch.cuda.empty_cache()
    sampler = ClockSampler(local)
    if rank == 0 and sample_clocks and not os.environ.get("HTCN_BENCH_NO_SAMPLER"):
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
    tr.allreduce_events = []
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
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 and sample_clocks else None
    ar = [a0.elapsed_time(a1) for a0, a1 in tr.allreduce_events]
    tr.allreduce_events = None
    t_ar = torch.tensor([float(np.mean(ar)) if ar else 0.0], device="cuda")
    if world > 1:
        dist.all_reduce(t_ar, op=dist.ReduceOp.MAX)
    allreduce_ms = float(t_ar.item())
    loss_dev = float(sc.cpu().numpy()[0])
    if world > 1:       # the replicas must hold bit-identical parameters after the timed steps
        cs = torch.stack([tr.params.double().sum(), tr.params.double().abs().sum()])
        cmax, cmin = cs.clone(), cs.clone()
        dist.all_reduce(cmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(cmin, op=dist.ReduceOp.MIN)
        if rank == 0:
            dp_parity["replicas_identical"] = bool(torch.equal(cmax, cmin))
            dp_parity["ok"] = dp_parity["ok"] and dp_parity["replicas_identical"]
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
    if rank != 0:
        return None
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
            "dp_parity": dp_parity,
            "allreduce": {"ms_per_step": allreduce_ms, "share_of_step": allreduce_ms / ms_per_step, "bytes": int(tr.n_flat) * 4,
                          "path": ("one peer-memory kernel: reduce-scatter + all-gather of the flat gradient over NVLink, the scalar "
                                   "all-reduce and the Adam update fused (htcn_peer_allreduce_adam); ms_per_step covers all of it")
                                  if tr.peer is not None else
                                  "NCCL all-reduce of the flat gradient buffer + of the 8 scalars between backward and Adam (Adam not included)",
                          "overlap": "none (device time between two events on the compute stream, max over ranks)"},
            "roofline": {"kernel": "whole step, catalog products only", "bound": "tensor", "achieved": flops / (ms_per_step * 1e-3) / 1e12 / world,
                         "peak": peaks["tf_sust"], "unit": "TFLOP/s", "frac": flops / (ms_per_step * 1e-3) / 1e12 / world / peaks["tf_sust"],
                         "peak_source": peaks["src"], "traffic": None},
            "clocks": clocks}
    return line


def sharded_leg(opt, rank, world, local, items=8 * 1024 * 1024, queries=4096, k=100, parity_items=1_000_000):
    """BASELINE configs[3]: CE + rank + top-k of `queries` query rows over `items` items, the catalog sharded over the
    ranks (rank r owns rows [n0_r, n1_r) of W_out^T in the bf16 scoring layout).  Device time per phase / collective (max
    over ranks), and an in-run parity check of the sharded path against the replicated single-GPU path on a catalog of
    `parity_items` items: ranks, top-k indices and values bit for bit; loss rows to 1e-5 relative -- the CE partial sums are
    fp32 sums of a million exponentials taken in a different order (world x 3 splits against 4): measured 2.3e-5 / 2.7e-5 /
    3.2e-5 absolute at 2 / 4 / 8 shards on losses near 20 (profiles/r2_sharded_loss_order.txt); the maximum of this run is
    reported as `loss_max_abs_diff`."""
    import torch
    import torch.distributed as dist
    from hiertcn_b200 import _cabi as cabi
    from hiertcn_b200.dist import CatalogTable, CudaScoreOps, ShardedCatalogScorer, choose_n_split, shard_bounds
    from hiertcn_b200.peer import PeerShardedCatalogScorer
    dev = torch.device("cuda", local)
    cabi.load()

    def table(n_rows, seed):
        g = torch.Generator(device=dev).manual_seed(seed)
        wt = torch.empty((n_rows, cabi.WT_PITCH_BF16), dtype=torch.bfloat16, device=dev)
        step = 1 << 20                                   # prepare in slices: the fp32 TF-layout source is 4x the table
        for r0 in range(0, n_rows, step):
            n = min(step, n_rows - r0)
            w = torch.randn((128, n), device=dev, generator=g) * 0.3
            b = torch.randn(n, device=dev, generator=g) * 0.2
            cabi.call("htcn_prepare_wout", w.data_ptr(), b.data_ptr(), n, wt[r0:r0 + n].data_ptr(), cabi.HTCN_BF16,
                      torch.cuda.current_stream(dev).cuda_stream)
            torch.cuda.synchronize(dev)
        return wt

    def n_split_for(Q):
        return choose_n_split(Q, items // world, torch.cuda.get_device_properties(dev).multi_processor_count)

    Ql = queries // world
    # ---- parity: the same 1M-item table on every rank; sharded N ways vs scored whole on this rank
    full = CatalogTable(table(parity_items, 4242), None, "bf16", parity_items)
    gq = torch.Generator(device=dev).manual_seed(900 + rank)
    Qp = min(Ql, 512)
    hp = torch.randn((Qp, 128), device=dev, generator=gq).to(torch.bfloat16)
    yp = torch.randint(1, parity_items, (Qp,), device=dev, generator=gq, dtype=torch.int32)
    b = shard_bounds(parity_items, world)
    # the exchanges run as our own peer-memory kernels (hiertcn_b200.peer); the NCCL-collective path is timed beside them.
    # A peer-buffer setup failure (no CUDA IPC between the ranks) raises on every rank alike: fall back and say so.
    exchange = "nccl collectives"
    sc = ShardedCatalogScorer(CudaScoreOps(full.rows(b[rank], b[rank + 1])), dist, rank, world, parity_items, n_split=3)
    psc = None
    if world > 1 and not os.environ.get("HTCN_SHARDED_NCCL"):
        why = ""
        try:
            psc = PeerShardedCatalogScorer(sc.ops, dist, rank, world, parity_items, n_split=3)
            got = psc.score(hp, yp, k=k)
            torch.cuda.synchronize(dev)
            psc.check()
        except RuntimeError as e:
            why = str(e)[:120]
        fine = torch.tensor([0 if why else 1], device=dev)
        dist.all_reduce(fine, op=dist.ReduceOp.MIN)                  # one rank's failure sends every rank to the collectives
        if int(fine.item()):
            exchange = "peer-memory kernels (htcn_peer_exchange / htcn_peer_bcast_owned over NVLink, CUDA IPC)"
        else:
            exchange = "nccl collectives (peer path unavailable: %s)" % (why or "failed on another rank")
            if psc is not None:
                psc.close()
            psc = None
    if psc is None:
        got = sc.score(hp, yp, k=k)
    ops = CudaScoreOps(full)
    zy = torch.zeros(Qp, dtype=torch.float32, device=dev)
    ops.target_logit(hp, yp, 0, parity_items, zy)
    part = ops.sweep(hp, yp, zy, 0, parity_items, k, 4, True, True)
    ref = ops.finish(part["pm"], part["ps"], part["pc"], yp, zy)
    ref.update(ops.topk_merge(part["tv"], part["ti"], k))
    torch.cuda.synchronize(dev)
    checks = dict(rank=torch.equal(got["rank_row"], ref["rank_row"]), topk_idx=torch.equal(got["topk_idx"], ref["topk_idx"]),
                  topk_val=torch.equal(got["topk_val"], ref["topk_val"]),
                  loss=bool(torch.allclose(got["loss_row"], ref["loss_row"], rtol=1e-5, atol=1e-5)))
    ok = torch.tensor([int(all(checks.values()))] + [int(v) for v in checks.values()], device=dev)
    loss_diff = (got["loss_row"] - ref["loss_row"]).abs().max().reshape(1)
    if world > 1:
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        dist.all_reduce(loss_diff, op=dist.ReduceOp.MAX)
    loss_diff = float(loss_diff.item())
    if psc is not None:
        psc.close()
    del full, ops, part, ref, got, sc, psc
    torch.cuda.empty_cache()
    # ---- timing: 8M items, this rank's shard only
    b = shard_bounds(items, world)
    n0, n1 = b[rank], b[rank + 1]
    shard = CatalogTable(table(n1 - n0, 100 + rank), None, "bf16", items)
    h = torch.randn((Ql, 128), device=dev, generator=gq).to(torch.bfloat16)
    y = torch.randint(1, items, (Ql,), device=dev, generator=gq, dtype=torch.int32)
    ns = n_split_for(queries)
    names = ["allgather_queries", "target_logit", "allreduce_target", "sweep", "alltoall_partials", "merge"]
    steps = max(3, opt.steps)

    def timed(sc):
        for _ in range(3):
            out = sc.score(h, y, k=k)
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        sc.phases = []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            out = sc.score(h, y, k=k)
        e1.record()
        torch.cuda.synchronize(dev)
        phases = sc.phase_ms()
        sc.phases = None
        t = torch.tensor([e0.elapsed_time(e1) / steps] + [phases.get(n, 0.0) / steps for n in names], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t, out

    shard_ops = CudaScoreOps(shard)
    t_nccl = None
    if exchange.startswith("peer"):
        psc = PeerShardedCatalogScorer(shard_ops, dist, rank, world, items, n_split=ns)
        t, out = timed(psc)
        psc.check()
        t_nccl, out_nccl = timed(ShardedCatalogScorer(shard_ops, dist, rank, world, items, n_split=ns))
        same = torch.tensor([int(all(torch.equal(out[n], out_nccl[n]) for n in out_nccl))], device=dev)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        peer_equals_nccl = bool(same.item())
    else:
        t, out = timed(ShardedCatalogScorer(shard_ops, dist, rank, world, items, n_split=ns))
        peer_equals_nccl = None
    v, i = out["topk_val"], out["topk_idx"]
    sane = bool((v[:, :-1] >= v[:, 1:]).all() and (i >= 0).all() and (i < items).all() and torch.isfinite(out["loss_row"]).all())
    if exchange.startswith("peer"):
        psc.close()
    if rank != 0:
        return None
    ms = float(t[0].item())
    ph = {n: float(t[1 + j].item()) for j, n in enumerate(names)}
    coll = {n: ph[n] for n in ("allgather_queries", "allreduce_target", "alltoall_partials")}
    rec = {"workload": "CE + rank + top-%d of %d queries over %d items, catalog sharded %d ways (bf16 tier)" % (k, queries, items, world),
           "exchange": exchange,
           "ms_per_call": ms, "queries_per_s": queries / (ms * 1e-3), "n_split": ns,
           "useful_tflops_per_gpu": 2.0 * queries * 128 * (items / world) / (ms * 1e-3) / 1e12,
           "phase_ms": ph, "collective_ms": sum(coll.values()), "limiting_collective": max(coll, key=coll.get) if world > 1 else None,
           "sorted_in_range_finite": sane, "sharded_parity": bool(ok[0].item()),
           "parity_checks": {n: bool(ok[1 + j].item()) for j, n in enumerate(checks)}, "loss_max_abs_diff": loss_diff,
           "parity_config": "%d queries per rank x %d items: sharded %d ways vs whole catalog on one GPU" % (Qp, parity_items, world)}
    if t_nccl is not None:
        phn = {n: float(t_nccl[1 + j].item()) for j, n in enumerate(names)}
        rec["nccl_path"] = {"ms_per_call": float(t_nccl[0].item()), "phase_ms": phn,
                            "collective_ms": sum(phn[n] for n in coll), "same_results_as_peer_path": peer_equals_nccl}
    return rec


def kernels_leg(model, staged, neg_dev, peaks, wl):
    """Device time of every kernel group of one step on its own (CUDA events on the launching stream, inputs of the whole
    batch = far larger than L2), with the roofline each is held against (SURVEY 8d algorithmic bytes / FLOPs)."""
    import torch
    from hiertcn_b200 import _cabi as cabi
    B, T, S, Q = staged["B"], staged["T"], staged["S"], staged["Q"]
    st = model.stream_ptr()
    f32 = torch.float32
    slot_p, keep = cabi.int_array(staged["slot_off"])
    xe = model._buf("xe", (B * T, 128), model.act_torch_dtype)
    yp = model._buf("yp", (S, B, 128), f32)
    sbias = model._buf("sbias", (S, B, 128), f32)
    state_out = torch.empty((B, 256), dtype=f32, device=model.device)
    hout = model._buf("hout", (max(Q, 1), 128), model.act_torch_dtype)
    k3s = model._buf("k3_scratch", (cabi.gru_scratch_bytes(B) // 4,), f32)
    k2s = model._buf("k2_scratch_bf16", (cabi.tcn_scratch_floats(model.n_levels, model.K),), f32)
    g = model._gru_pp
    e = 2 if model.precision == "bf16" else 4

    def k1():
        cabi.call("htcn_gather_meanpool", model.E.data_ptr(), model.emb_pitch, model.b_emb.data_ptr(), model.N, staged["x_id"].data_ptr(),
                  staged["y_id"].data_ptr(), slot_p, B, T, S, xe.data_ptr(), model.act_dtype, yp.data_ptr(), st)

    def k3():
        cabi.call("htcn_gru_sessions", yp.data_ptr(), staged["mask"].data_ptr(), staged["state"].data_ptr(), g[0][0], g[1][0], g[2][0],
                  g[3][0], 2, model.w_in_state.data_ptr(), B, S, model.act_dtype, k3s.data_ptr(), None, sbias.data_ptr(),
                  state_out.data_ptr(), st)

    def k2():
        cabi.call("htcn_tcn_forward", xe.data_ptr(), model.act_dtype, model._k2_precision(), model.w_in_x.data_ptr(), sbias.data_ptr(),
                  model._conv_w_pp[0], model._conv_b_pp[0], None, None, model.n_levels, model.K, slot_p, B, T, S, staged["row_of"].data_ptr(),
                  hout.data_ptr(), model.act_dtype, k2s.data_ptr(), st)

    out = {}
    scores, _ = model.forward(staged=staged)

    def sl():
        model.sampled_loss(scores, neg_dev)

    row_b = model.emb_pitch * 4                              # bytes of one gathered table row (packed: emb_dim floats)
    works = {
        "k1_gather_meanpool": (k1, "hbm", B * (2 * T * row_b + 2 * T * 4 + T * 128 * e + S * 512),
                               "rows read for x and y (%d B each, packed emb_dim) + ids + Xe/Yp written" % row_b),
        "k3_gru_sessions": (k3, "hbm", B * S * (128 + 2 * 256) * 4, "Yp in, state_pre/sbias out; latency-bound in practice"),
        "k2_tcn_forward": (k2, "tensor", Q * (model.n_levels * 2.0 * model.K * 128 * 128 + 2.0 * 128 * 128), "useful FLOPs of the scored positions"),
        "sampled_rank_loss": (sl, "hbm", Q * (21 * 512 + 128 * e + 21 * 4 + 4), "21 gathered fp32 rows of W_out^T per position + the query row"),
    }
    reps = 10
    for name, (fn, bound, work, note) in works.items():
        # sub-millisecond kernels timed right after an idle gap run at ramping clocks: warm up first, no host sync between
        # the warm-up and the timed calls
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(model.device)
        for _ in range(10):
            fn()
        e0.record(torch.cuda.current_stream(model.device))
        for _ in range(reps):
            fn()
        e1.record(torch.cuda.current_stream(model.device))
        torch.cuda.synchronize(model.device)
        ms = e0.elapsed_time(e1) / reps
        peak = peaks["hbm"] if bound == "hbm" else peaks["tf_burst"]
        ach = work / (ms * 1e-3) / (1e9 if bound == "hbm" else 1e12)
        out[name] = {"ms": ms, "bound": bound, "achieved": ach, "peak": peak, "unit": "GB/s" if bound == "hbm" else "TFLOP/s",
                     "frac": ach / peak, "work": note}
    del keep
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="bf16", choices=["bf16", "f32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-legs", action="store_true", help="skip the strong / sharded / train / kernels legs")
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
    # the fetch list of run_hier_xing.py:145-149: loss, state, the metric means and the per-position ranks_float map
    # (mask_y = sign(y_id) is host data already)
    fetch = ("ranks_float",)
    for _ in range(2):
        out = model.step(x, y, m, s0, neg_ids=neg_host, per_position=fetch)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e2e_steps = max(2, min(opt.steps, 5))
    t0 = time.perf_counter()
    pend = None
    for _ in range(e2e_steps):                  # depth-2 software pipeline: pack + H2D of step i+1 overlap step i
        nxt = model.step_async(x, y, m, s0, neg_ids=neg_host, per_position=fetch)
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
    d2h = 8 * 4 + B * 256 * 4 + B * T * 4 + 8 * 4          # scalars, state, ranks_float [B,T], sampled-loss scalars
    assert out["ranks_float"].shape == (B, T)

    # ---------------- extra legs (see module docstring) ----------------
    legs = {}
    if not opt.no_legs:
        # a leg that fails (on every rank alike: a shape that does not divide, an allocation) is reported in its place and
        # must not take the headline line with it
        def guarded(name, fn):
            try:
                return fn()
            except Exception as e:      # noqa: BLE001
                import traceback
                traceback.print_exc()
                return {"error": "%s leg failed: %s" % (name, repr(e)[:300])} if rank == 0 else None

        kern = guarded("kernels", lambda: kernels_leg(model, staged, neg_dev, peaks, wl)) if rank == 0 else None
        # strong scaling of configs[1]: the global batch of wl["B"] users split over the ranks
        if world > 1 and wl["B"] % world == 0:
            Bs = wl["B"] // world
            xs, ys, ms_, ss = make_inputs(wl, seed=50 + rank, B=Bs)
            staged_s = model.stage(xs, ys, ms_, ss)
            neg_s = neg_dev[:int(staged_s["Q"])]

            def strong_step():
                scores, _ = model.forward(staged=staged_s)
                model.sampled_loss_mean(scores, neg_s)
                return reduce_scalars(model.loss(scores, metrics=True)["scalars"])

            for _ in range(3):
                strong_step()
            drain()
            torch.cuda.synchronize()
            dist.barrier()
            s0e, s1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0e.record()
            for _ in range(opt.steps):
                strong_step()
            drain()
            s1e.record()
            torch.cuda.synchronize()
            ts = torch.tensor([s0e.elapsed_time(s1e) / opt.steps], device="cuda")
            dist.all_reduce(ts, op=dist.ReduceOp.MAX)
            legs["strong"] = {"workload": "configs[1], global batch %d users split over %d GPUs" % (wl["B"], world),
                              "users_per_gpu": Bs, "ms_per_step": float(ts.item()), "value": wl["B"] / (float(ts.item()) * 1e-3),
                              "unit": UNIT, "scaling": "strong"}
            del staged_s
        else:
            legs["strong"] = {"workload": "configs[1], global batch %d users on 1 GPU" % wl["B"], "users_per_gpu": wl["B"],
                              "ms_per_step": ms_per_step, "value": value, "unit": UNIT, "scaling": "strong"}
        if kern is not None:
            legs["kernels"] = kern
        sh = guarded("sharded", lambda: sharded_leg(opt, rank, world, local))
        if sh is not None:
            legs["sharded"] = sh
        torch.cuda.empty_cache()
        tl = guarded("train", lambda: train_leg(opt, dict(WORKLOADS["cfg5"]), rank, world, local))
        if tl is not None and "error" in tl:
            legs["train"] = tl
        elif tl is not None:
            legs["train"] = {k: tl[k] for k in ("metric", "value", "unit", "ms_per_step", "scaling", "config", "loss", "e2e",
                                                "gpu_launches", "allreduce", "dp_parity", "roofline")}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---------------- roofline of the dominant kernel (K4 sweep) ----------------
    sweep_ms = float(np.mean([s[0] for s in sweep]))
    achieved = sweep[0][1] / (sweep_ms * 1e-3) / 1e12
    peak = peaks["tf_sust"]
    roofline = {"kernel": "k4_score_bf16_cg2<CE|RANK, target folded into the MMA, packed f32x2 epilogue>" if opt.precision == "bf16" else "k4_score_f32",
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
    line.update(legs)
    if world > 1:
        dist.destroy_process_group()
    if not opt.no_cpu_baseline and world == 1:          # the CPU baseline is reported at N=1 only
        base, _, _, _ = cpu_baseline(wl, w, seconds_target=12.0)
        base["literal_cfg1"] = cpu_literal_cfg1()
        line["cpu_baseline"] = base
    print(json.dumps(line))


if __name__ == "__main__":
    main()
