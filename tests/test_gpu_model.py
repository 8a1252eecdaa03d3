"""GPU parity of the whole hot path through the Python surface (HierTCN.step == the reference's
sess.run fetch list) against the golden fixtures and the CPU oracle."""
import numpy as np
import pytest

from helpers import golden_variant_kwargs, load_hier_golden, small_case
from oracle import hiertcn_oracle as O

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def make_model(w, N, precision, levels=2, K=5):
    from hiertcn_b200.args import make_args
    from hiertcn_b200.model_hier import HierTCN
    a = make_args(["--item_num", str(N), "--tcn_channel", ",".join(["128"] * levels), "--kernel_size", str(K)])
    return HierTCN(a, w, precision=precision).build()


def test_step_f32_matches_golden_reference_run():
    """fixture produced by running the reference's own model_hier/loss python (oracle/make_golden.py)"""
    z, x, y, m, w = load_hier_golden("hier_default_arch")
    model = make_model(w, int(z["N"]), "f32")
    out = model.step(x, y, m, z["state0"], per_position=True)
    assert abs(out["loss"] - z["loss_f64"]) <= 1e-4 * abs(z["loss_f64"])
    np.testing.assert_allclose(out["state"], z["state_f64"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(out["loss_bt"], z["loss_bt_f64"], rtol=1e-4, atol=1e-5)
    np.testing.assert_array_equal(out["ranks"], z["ranks_f64"])      # N=61, well separated logits
    got = np.asarray([out[k] for k in ("recall1", "recall5", "recall10", "mrr", "mrp")])
    np.testing.assert_allclose(got, z["metrics_f64"], rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("B,S,L,N,lengths", [(64, 10, 20, 20778, "ragged"), (9, 4, 20, 997, "dense"), (3, 2, 1, 40, "dense")])
def test_step_f32_matches_oracle(B, S, L, N, lengths):
    x, y, m, s0, w = small_case(B=B, S=S, L=L, N=N, seed=B, lengths=lengths)
    ref = O.forward_loss_metrics(x, y, m, s0, w, 2, "f64")
    model = make_model(w, N, "f32")
    out = model.step(x, y, m, s0, per_position=True, topk=min(100, N))
    assert abs(out["loss"] - ref["loss"]) <= 1e-4 * abs(ref["loss"])
    np.testing.assert_allclose(out["state"], ref["state"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(out["loss_bt"], ref["loss_bt"], rtol=1e-4, atol=2e-5)
    y_id = np.concatenate([np.asarray(v) for v in y], 1).astype(np.int64)
    amb = O.rank_ambiguity(ref["pred"], y_id, 2e-5) * (y_id > 0)
    assert (np.abs(out["ranks"] - ref["ranks"]) <= amb).all()
    assert np.mean(out["ranks"] == ref["ranks"]) > 0.98
    assert abs(out["mrr"] - ref["mrr"]) < 2e-3 and abs(out["mrp"] - ref["mrp"]) < 1e-4
    # top-k sets of the valid rows vs the oracle's logits: identical wherever the k-th gap is not a near-tie
    valid = y_id.reshape(-1) > 0
    zv = ref["pred"].reshape(-1, N)[valid]
    k = out["topk_idx"].shape[1]
    v_ref, i_ref = O.top_k(zv, k)
    same = np.array([set(a) == set(b) for a, b in zip(out["topk_idx"], i_ref)])
    srt = -np.sort(-zv, axis=1)
    gap = (srt[:, k - 1] - srt[:, k]) if N > k else np.ones(len(zv))
    assert same[gap > 1e-4].all()
    np.testing.assert_allclose(out["topk_val"], v_ref, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("precision,tol", [("f32", 1e-4), ("bf16", 2e-2)])
def test_step_matches_golden_downsample_narrow_widths(precision, tol):
    """fixture produced by the reference's own python with tcn_channel = [32, 32, 48], hidden_dim = 16, kernel_size = 3:
    two levels use the 1x1 down-sample residual (customized_tcn_cell.py:102-106,123-126); every width runs zero-padded"""
    from hiertcn_b200.args import make_args
    from hiertcn_b200.model_hier import HierTCN
    z, x, y, m, w = load_hier_golden("hier_downsample_3lvl")
    N = int(z["N"])
    a = make_args(["--item_num", str(N), "--tcn_channel", "32,32,48", "--kernel_size", "3", "--hidden_dim", "16"])
    model = HierTCN(a, w, precision=precision).build()
    out = model.step(x, y, m, z["state0"], per_position=True)
    assert out["state"].shape == z["state_f64"].shape
    assert abs(out["loss"] - z["loss_f64"]) <= tol * abs(z["loss_f64"])
    np.testing.assert_allclose(out["state"], z["state_f64"], rtol=max(tol, 1e-4), atol=max(tol, 1e-5))
    np.testing.assert_allclose(out["loss_bt"], z["loss_bt_f64"], rtol=tol, atol=max(tol, 1e-5))
    if precision == "f32":
        np.testing.assert_array_equal(out["ranks"], z["ranks_f64"])
        got = np.asarray([out[k] for k in ("recall1", "recall5", "recall10", "mrr", "mrp")])
        np.testing.assert_allclose(got, z["metrics_f64"], rtol=1e-4, atol=1e-6)
        pred = model.forward(x, y, m, z["state0"])[0].materialize()
        np.testing.assert_allclose(pred, z["pred_f64"], rtol=1e-4, atol=1e-4)
    else:
        assert abs(out["mrr"] - z["metrics_f64"][3]) < 5e-2
    # the carried state round-trips through the host in the TF shape [B, G*H]
    out2 = model.step(x, y, m, out["state"])
    assert np.isfinite(out2["loss"])


@pytest.mark.parametrize("precision,tol", [("f32", 1e-4), ("bf16", 2e-2)])
def test_step_matches_golden_gap_decay_l2norm_warmstart(precision, tol):
    """fixture produced by the reference's own python with has_gap (model_hier.py:40-47), l2_normalize (model_tcn.py:42-43)
    and a warm-start loss mask (model.py:102-103)"""
    from hiertcn_b200.args import make_args
    from hiertcn_b200.model_hier import HierTCN
    z, x, y, m, w = load_hier_golden("hier_gap_l2norm_warmstart")
    kw = golden_variant_kwargs(z)
    N = int(z["N"])
    a = make_args(["--item_num", str(N), "--has_gap", "--l2_normalize", "--gap_bandwidth", str(kw["gap_bandwidth"])])
    model = HierTCN(a, w, precision=precision).build()
    out = model.step(x, y, m, z["state0"], per_position=True, topk=10, mask_warmstart=kw["mask_warmstart"], x_gap=kw["x_gap"])
    assert abs(out["loss"] - z["loss_f64"]) <= tol * abs(z["loss_f64"])
    np.testing.assert_allclose(out["state"], z["state_f64"], rtol=max(tol, 1e-4), atol=max(tol, 1e-5))
    np.testing.assert_allclose(out["loss_bt"], z["loss_bt_f64"], rtol=tol, atol=max(tol, 1e-5) * 0.1)
    if precision == "f32":
        np.testing.assert_array_equal(out["ranks"], z["ranks_f64"])
        got = np.asarray([out[k] for k in ("recall1", "recall5", "recall10", "mrr", "mrp")])
        np.testing.assert_allclose(got, z["metrics_f64"], rtol=1e-4, atol=1e-6)
        # the materialised logits are the NORMALISED ones; the fixture's are additionally zeroed by the warm-start mask
        scores, _ = model.forward(x, y, m, z["state0"], x_gap=kw["x_gap"])
        pred = scores.materialize()
        y_id = np.concatenate(y, 1)
        keep = (y_id > 0) & (kw["mask_warmstart"] > 0)
        np.testing.assert_allclose(pred[keep], z["pred_f64"][keep], rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(np.linalg.norm(pred[y_id > 0], axis=-1), 1.0, rtol=1e-5)
        # top-k: normalised values in the reference's order
        ref = O.forward_loss_metrics(x, y, m, z["state0"], w, 2, "f64", **kw)
        rows = out["row_of"].reshape(y_id.shape)
        zv = ref["pred"][keep]
        v_ref, i_ref = O.top_k(zv, 10)
        np.testing.assert_allclose(out["topk_val"][rows[keep]], v_ref, rtol=1e-4, atol=1e-6)
        assert np.mean(out["topk_idx"][rows[keep]] == i_ref) > 0.99
    else:
        assert abs(out["mrr"] - z["metrics_f64"][3]) < 5e-2


def test_materialized_logits_match_reference_pred():
    z, x, y, m, w = load_hier_golden("hier_default_arch")
    from hiertcn_b200.args import make_args
    from hiertcn_b200.model_hier import model_hier
    a = make_args(["--item_num", str(int(z["N"]))])
    pred, state = model_hier(a, x, y, m, z["state0"], weights=w, precision="f32")
    np.testing.assert_allclose(pred.materialize(), z["pred_f64"], rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(state, z["state_f64"], rtol=1e-4, atol=1e-5)


def test_state_carry_across_batches_and_user_independence():
    x, y, m, s0, w = small_case(B=8, S=3, L=6, N=101, seed=21)
    model = make_model(w, 101, "f32")
    o1 = model.step(x, y, m, s0, per_position=True)
    perm = np.random.default_rng(0).permutation(8)
    o2 = model.step([a[perm] for a in x], [a[perm] for a in y], [a[perm] for a in m], s0[perm], per_position=True)
    np.testing.assert_array_equal(o1["loss_bt"][perm], o2["loss_bt"])       # rows are independent users
    np.testing.assert_array_equal(o1["state"][perm], o2["state"])
    o3 = model.step(x, y, m, o1["state"])                                     # carried state feeds the next batch
    o1d = model.step(x, y, m, s0, state_on_device=True)                       # ... or stays on the device
    assert o1d["state"].is_cuda and np.array_equal(o1d["state"].cpu().numpy(), o1["state"])
    o3d = model.step(x, y, m, o1d["state"])
    assert o3d["loss"] == o3["loss"]
    ref = O.forward_loss_metrics(x, y, m, o1["state"], w, 2, "f64")
    assert abs(o3["loss"] - ref["loss"]) <= 1e-4 * abs(ref["loss"])


def test_bf16_tier_within_2e2():
    x, y, m, s0, w = small_case(B=40, S=5, L=12, N=5000, seed=33)
    ref = O.forward_loss_metrics(x, y, m, s0, w, 2, "f64")
    model = make_model(w, 5000, "bf16")
    out = model.step(x, y, m, s0, per_position=True, topk=100)
    assert abs(out["loss"] - ref["loss"]) <= 2e-2 * abs(ref["loss"])
    np.testing.assert_allclose(out["loss_bt"], ref["loss_bt"], rtol=2e-2, atol=2e-2)
    np.testing.assert_allclose(out["state"], ref["state"], rtol=2e-2, atol=2e-2)     # tensor-core GRU: bf16 operands, fp32 state
    model.k3_tcgen05 = False                                                          # the fp32 recurrence is still selectable
    out32 = model.step(x, y, m, s0)
    np.testing.assert_allclose(out32["state"], ref["state"], rtol=1e-4, atol=1e-5)
    model.k3_tcgen05 = True
    assert abs(out["mrr"] - ref["mrr"]) < 2e-2 and abs(out["mrp"] - ref["mrp"]) < 2e-3
    # top-k sets vs the bf16-emulating oracle: high overlap (operands rounded identically, order differs)
    refb = O.forward_loss_metrics(x, y, m, s0, w, 2, "bf16")
    y_id = np.concatenate([np.asarray(v) for v in y], 1).astype(np.int64)
    zv = refb["pred"].reshape(-1, 5000)[y_id.reshape(-1) > 0]
    _, i_ref = O.top_k(zv, 100)
    overlap = np.mean([len(set(a) & set(b)) / 100.0 for a, b in zip(out["topk_idx"], i_ref)])
    assert overlap > 0.97, overlap


def test_step_bf16_exact_baseline_config1():
    """BASELINE.json configs[0] exactly (batch 64, 10 sessions x 20 positions, 20 778 items) in the bf16 tier, end to end
    against the fp64 oracle at the north star's 2e-2 bar"""
    B, S, L, N = 64, 10, 20, 20778
    x, y, m, s0, w = small_case(B=B, S=S, L=L, N=N, seed=64, lengths="dense")
    ref = O.forward_loss_metrics(x, y, m, s0, w, 2, "f64")
    model = make_model(w, N, "bf16")
    out = model.step(x, y, m, s0, per_position=True)
    assert abs(out["loss"] - ref["loss"]) <= 2e-2 * abs(ref["loss"])
    np.testing.assert_allclose(out["loss_bt"], ref["loss_bt"], rtol=2e-2, atol=2e-2)
    np.testing.assert_allclose(out["state"], ref["state"], rtol=2e-2, atol=2e-2)
    assert abs(out["mrr"] - ref["mrr"]) < 2e-2 and abs(out["mrp"] - ref["mrp"]) < 2e-3
    for kk in ("recall1", "recall5", "recall10"):
        assert abs(out[kk] - ref[kk]) < 2e-2
    # ranks: integer counts on bf16-rounded operands; against the fp64 logits they may differ by the number of columns
    # within the bf16 tier's logit error of the target
    y_id = np.concatenate([np.asarray(v) for v in y], 1).astype(np.int64)
    amb = O.rank_ambiguity(ref["pred"], y_id, 2e-2) * (y_id > 0)
    assert np.mean(np.abs(out["ranks"] - ref["ranks"]) <= amb) > 0.99
    assert np.abs(out["ranks_float"] - ref["ranks_float"]).max() < 2e-2


def test_evaluate_hier_loop_matches_oracle_batch_by_batch():
    """queue loader (reference enqueue/dequeue semantics) -> evaluate_hier with device-resident carried state ==
    the oracle fed the same batches with the state carried on the host (run_hier_xing.py:83-207)"""
    from hiertcn_b200.args import make_args
    from hiertcn_b200.data_loader import Dataloader_hier_model_xing, make_synthetic_interactions
    from hiertcn_b200.model_hier import HierTCN
    from hiertcn_b200.run_hier import evaluate_hier
    from hiertcn_b200.weights import hier_weight_shapes, init_weights
    N = 300
    table, data = make_synthetic_interactions(120, N, seed=4)
    a = make_args(["--item_num", str(N), "--batch_size", "6", "--max_session_num", "4", "--max_activity_len", "8"])
    w = init_weights(hier_weight_shapes(N), seed=8, kernel_scale=2.0, bias_noise=0.1)
    model = HierTCN(a, w, precision="f32").build()
    res = evaluate_hier(model, Dataloader_hier_model_xing(a, "train", data=(table, data)), max_batches=4)
    ld = Dataloader_hier_model_xing(a, "train", data=(table, data))
    state = np.zeros((6, 256), np.float32)
    acc = dict(loss=0.0, mrr=0.0, mrp=0.0, recall10=0.0)
    for _ in range(4):
        x, y, m, _ = ld.get_batch()
        ref = O.forward_loss_metrics(x, y, m, state, w, 2, "f64")
        state = ref["state"]
        for k in acc:
            acc[k] += float(ref[k])
    assert res["batches"] == 4
    for k in acc:
        assert abs(res[k] - acc[k] / 4) <= 1e-4 * max(1.0, abs(acc[k] / 4)), (k, res[k], acc[k] / 4)
    assert len(res["rank_by_position"]) >= 4 and 0.0 <= res["user_rank_mean"] <= 1.0


def test_loss_module_surface():
    """hiertcn_b200.loss mirrors reference loss.py on the lazy scores handle"""
    from hiertcn_b200 import loss as L
    x, y, m, s0, w = small_case(B=6, S=3, L=7, N=211, seed=9)
    ref = O.forward_loss_metrics(x, y, m, s0, w, 2, "f64")
    model = make_model(w, 211, "f32")
    scores, _ = model.forward(x, y, m, s0)
    np.testing.assert_allclose(L.calc_loss(scores).cpu().numpy(), ref["loss_bt"], rtol=1e-4, atol=2e-5)
    rec1, rec5, rec10, mrr, mrp, ranks_float, ranks = L.calc_metric_fast(scores)
    np.testing.assert_array_equal(ranks.cpu().numpy(), ref["ranks"])
    assert abs(float(mrr) - ref["mrr"]) < 1e-5 and abs(float(rec10) - ref["recall10"]) < 1e-6
    # calc_score against the oracle on the user embeddings the kernels produced
    rng = np.random.default_rng(0)
    cand = rng.integers(1, 211, size=(scores.Q, 5)).astype(np.int32)
    hout = scores.hout.float().cpu().numpy()
    table = w["hier/tcn/dense/kernel"].T
    for mode in ("l2", "inner_prod"):
        got = L.calc_score(scores, cand, mode).cpu().numpy()
        want = O.calc_score(hout[None].astype(np.float64), table[cand][None].astype(np.float64), mode)[0]
        np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-4)
    v, i = L.top_k(scores, 10)
    z = (hout.astype(np.float64) @ w["hier/tcn/dense/kernel"].astype(np.float64) + w["hier/tcn/dense/bias"])
    _, i_ref = O.top_k(z, 10)
    assert np.mean(i.cpu().numpy() == i_ref) > 0.99


@pytest.mark.parametrize("seed", list(range(8)))
def test_randomised_shapes_both_tiers(seed):
    """random batch geometry / catalog size / conv stack: fp32 tier within 1e-4, bf16 tier within 2e-2 of the fp64 oracle"""
    rng = np.random.default_rng(1000 + seed)
    B, S, L = int(rng.integers(1, 70)), int(rng.integers(1, 12)), int(rng.integers(1, 22))
    N = int(rng.choice([2, 17, 129, 257, 1000, 3001]))
    levels, K = int(rng.integers(1, 4)), int(rng.integers(2, 6))
    x, y, m, s0, w = small_case(B=B, S=S, L=L, N=N, seed=seed, tcn_channel=(128,) * levels, kernel_size=K,
                                lengths="ragged" if seed % 2 else "dense", kernel_scale=1.5)
    ref = O.forward_loss_metrics(x, y, m, s0, w, 2, "f64")
    y_id = np.concatenate([np.asarray(v) for v in y], 1).astype(np.int64)
    for precision, tol in (("f32", 1e-4), ("bf16", 2e-2)):
        model = make_model(w, N, precision, levels=levels, K=K)
        out = model.step(x, y, m, s0, per_position=True, topk=min(5, N))
        assert abs(out["loss"] - ref["loss"]) <= tol * max(abs(ref["loss"]), 1e-3), (precision, out["loss"], ref["loss"])
        np.testing.assert_allclose(out["loss_bt"], ref["loss_bt"], rtol=tol, atol=tol)
        np.testing.assert_allclose(out["state"], ref["state"], rtol=max(tol, 1e-4), atol=max(tol, 1e-5))
        if precision == "f32":
            amb = O.rank_ambiguity(ref["pred"], y_id, 2e-5) * (y_id > 0)
            assert (np.abs(out["ranks"] - ref["ranks"]) <= amb).all()
        assert out["n_valid"] == (y_id > 0).sum() and out["user_count"] == ((y_id > 0).sum(1) > 0).sum()
        assert out["topk_idx"].shape == (int((y_id > 0).sum()), min(5, N))


@pytest.mark.parametrize("B,L,levels,precision", [(5, 40, 2, "f32"), (3, 300, 4, "f32"), (6, 300, 4, "bf16"), (9, 17, 3, "bf16")])
def test_single_level_model_tcn(B, L, levels, precision):
    """reference model_tcn.py (single-level TCN on the one-hot item sequence, BASELINE config 3 shape)"""
    from hiertcn_b200.args import make_args
    from hiertcn_b200.model_tcn import TCN
    from hiertcn_b200.weights import init_weights, tcn_weight_shapes
    N, K = 257, 5
    w = init_weights(tcn_weight_shapes(N, (128,) * levels, K, output_dim=N), seed=3, kernel_scale=1.5, bias_noise=0.1)
    rng = np.random.default_rng(B + L)
    x = rng.integers(1, N, size=(B, L))
    x[:, 0] = 0
    n = rng.integers(1, L + 1, size=B)
    y = np.where(np.arange(L)[None, :] < n[:, None], rng.integers(1, N, size=(B, L)), 0)
    a = make_args(["--item_num", str(N), "--tcn_channel", ",".join(["128"] * levels), "--kernel_size", str(K)])
    model = TCN(a, w, precision=precision).build()
    scores = model.forward(x, y)
    pred64 = O.model_tcn(O.one_hot_signed(x, N, np.float64), {k: v.astype(np.float64) for k, v in w.items()}, "tcn", "f64")
    loss, loss_bt, mask_y, act, uc, pred_m = O.hier_loss(pred64, y)
    met = O.calc_metric_fast(pred_m, mask_y, act, uc, y)
    tol = 1e-4 if precision == "f32" else 2e-2
    if precision == "f32":
        np.testing.assert_allclose(scores.materialize(), pred_m, rtol=1e-4, atol=2e-4)
    r = model.loss(scores, metrics=True, per_position=True)
    sc = r["scalars"].cpu().numpy()
    assert abs(sc[0] - loss) <= tol * abs(loss)
    np.testing.assert_allclose(r["loss_bt"].cpu().numpy(), loss_bt, rtol=tol, atol=tol)
    if precision == "f32":
        amb = O.rank_ambiguity(pred_m, y, 2e-5) * (y > 0)
        assert (np.abs(r["ranks"].cpu().numpy() - met[6]) <= amb).all()


@pytest.mark.parametrize("channels,K,L", [((128, 128, 128, 128, 256, 256), 5, 150), ((200, 256, 96), 3, 33), ((128, 256, 128), 5, 40)])
def test_single_level_tcn_with_256_channel_levels(channels, K, L):
    """the reference's single-level default stack [128,128,128,128,256,256] (args.py:310-311; dilations up to 32) and other
    stacks with 129..256-channel levels: two 128-wide planes on the fp32 kernels (down-sample residual where the width
    changes, K = 256 contraction in the scoring head), against the fp64 oracle -- logits, CE, ranks, top-k"""
    from hiertcn_b200.args import make_args
    from hiertcn_b200.model_tcn import TCN
    from hiertcn_b200.weights import init_weights, tcn_weight_shapes
    N, B = 301, 4
    w = init_weights(tcn_weight_shapes(N, channels, K, output_dim=N), seed=5, kernel_scale=1.5, bias_noise=0.1)
    rng = np.random.default_rng(L)
    x = rng.integers(1, N, size=(B, L))
    x[:, 0] = 0
    n = rng.integers(1, L + 1, size=B)
    y = np.where(np.arange(L)[None, :] < n[:, None], rng.integers(1, N, size=(B, L)), 0)
    a = make_args(["--item_num", str(N), "--tcn_channel", ",".join(str(c) for c in channels), "--kernel_size", str(K)])
    with pytest.raises(NotImplementedError):
        TCN(a, w, precision="bf16")
    model = TCN(a, w, precision="f32").build()
    scores = model.forward(x, y)
    pred64 = O.model_tcn(O.one_hot_signed(x, N, np.float64), {k: v.astype(np.float64) for k, v in w.items()}, "tcn", "f64")
    loss, loss_bt, mask_y, act, uc, pred_m = O.hier_loss(pred64, y)
    met = O.calc_metric_fast(pred_m, mask_y, act, uc, y)
    scale = np.abs(pred_m).max()
    np.testing.assert_allclose(scores.materialize(), pred_m, rtol=1e-4, atol=1e-4 * scale)
    r = model.loss(scores, metrics=True, per_position=True)
    sc = r["scalars"].cpu().numpy()
    assert abs(sc[0] - loss) <= 1e-4 * abs(loss)
    np.testing.assert_allclose(r["loss_bt"].cpu().numpy(), loss_bt, rtol=1e-4, atol=1e-4)
    amb = O.rank_ambiguity(pred_m, y, 2e-5) * (y > 0)
    assert (np.abs(r["ranks"].cpu().numpy() - met[6]) <= amb).all()
    tk = model.score(scores, ce=False, rank=False, topk=20)
    zv = pred64.reshape(-1, N)[y.reshape(-1) > 0]
    v_ref, i_ref = O.top_k(zv, 20)
    np.testing.assert_allclose(tk["topk_val"].cpu().numpy(), v_ref, rtol=1e-4, atol=1e-4 * scale)
    assert np.mean(tk["topk_idx"].cpu().numpy() == i_ref) > 0.98


def test_hier_with_a_256_channel_level():
    """model_hier with tcn_channel = [128, 256] (fp32 tier): conditioned TCN + GRU + 256-wide scoring head vs the oracle"""
    from hiertcn_b200.args import make_args
    from hiertcn_b200.model_hier import HierTCN
    x, y, m, s0, w = small_case(B=6, S=3, L=9, N=211, seed=31, tcn_channel=(128, 256), kernel_size=5)
    ref = O.forward_loss_metrics(x, y, m, s0, w, 2, "f64")
    a = make_args(["--item_num", "211", "--tcn_channel", "128,256"])
    model = HierTCN(a, w, precision="f32").build()
    out = model.step(x, y, m, s0, per_position=True, topk=10)
    assert abs(out["loss"] - ref["loss"]) <= 1e-4 * abs(ref["loss"])
    np.testing.assert_allclose(out["state"], ref["state"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(out["loss_bt"], ref["loss_bt"], rtol=1e-4, atol=2e-5)
    y_id = np.concatenate(y, 1).astype(np.int64)
    amb = O.rank_ambiguity(ref["pred"], y_id, 2e-5) * (y_id > 0)
    assert (np.abs(out["ranks"] - ref["ranks"]) <= amb).all()


@pytest.mark.parametrize("precision", ["f32", "bf16"])
def test_device_batcher_equals_host_loader(precision):
    """batches assembled in HBM (htcn_assemble_batch, fixed slot width) == the queue loader's host batches: same loss,
    metrics, carried state, and the same per-position ranks on the positions both layouts hold"""
    from hiertcn_b200.args import make_args
    from hiertcn_b200.data_loader import Dataloader_hier_model_xing, make_synthetic_interactions
    from hiertcn_b200.device_batcher import DeviceBatcher
    from hiertcn_b200.model_hier import HierTCN
    from hiertcn_b200.run_hier import evaluate_hier
    N = 997
    a = make_args(["--item_num", str(N), "--batch_size", "12", "--max_session_num", "4", "--max_activity_len", "9"])
    data = make_synthetic_interactions(80, N, seed=4)
    model = HierTCN(a, None, precision=precision).build()
    host = Dataloader_hier_model_xing(a, "train", data=data)
    devb = DeviceBatcher(a, data, passes=1)
    assert devb.n_batches >= 3
    st_h = st_d = None
    for k in range(3):
        x, y, m, _ = host.get_batch()
        oh = model.step(x, y, m, st_h, per_position=True, state_on_device=True)
        st_h = oh["state"]
        staged = devb.next(st_d)
        od = model.step(staged=staged, per_position=True, state_on_device=True)
        st_d = od["state"]
        for key in ("loss", "recall1", "recall5", "recall10", "mrr", "mrp", "user_count", "n_valid"):
            assert abs(oh[key] - od[key]) <= 1e-5 * max(1.0, abs(oh[key])), (k, key, oh[key], od[key])
        assert torch.allclose(st_h, st_d, rtol=1e-5, atol=1e-6)
        # host layout: slot s has width L_s = batch max; device layout: width 9 -> compare slot by slot
        off, L = 0, 9
        for s in range(4):
            w = y[s].shape[1]
            np.testing.assert_array_equal(oh["ranks"][:, off:off + w], od["ranks"][:, s * L:s * L + w])
            assert (od["ranks"][:, s * L + w:(s + 1) * L] == 0).all()
            off += w
    res = evaluate_hier(model, DeviceBatcher(a, data, passes=1), max_batches=2)
    assert res["batches"] == 2 and np.isfinite(res["loss"])
    # a FRESH model whose first call is the staged path (no host batch ever staged through it): the pinned result sets
    # must exist without stage() having run
    fresh = HierTCN(a, None, precision=precision).build()
    res2 = evaluate_hier(fresh, DeviceBatcher(a, data, passes=1), max_batches=2)
    assert res2["batches"] == 2 and abs(res2["loss"] - res["loss"]) <= 1e-6 * abs(res["loss"])


def test_stale_scores_handle_raises_and_shard_model_refuses_full_scoring():
    from hiertcn_b200 import _cabi as cabi
    from hiertcn_b200.args import make_args
    from hiertcn_b200.dist import make_sharded_model
    from hiertcn_b200.model_hier import HierTCN
    x, y, m, s0, w = small_case(B=4, S=2, L=5, N=600, seed=2)
    a = make_args(["--item_num", "600"])
    model = HierTCN(a, w, precision="f32").build()
    sc1, _ = model.forward(x, y, m, s0)
    l1 = float(model.loss(sc1)["scalars"][0].item())
    sc2, _ = model.forward([v[::-1].copy() for v in x], [v[::-1].copy() for v in y], [v[::-1].copy() for v in m], s0[::-1].copy())
    with pytest.raises(cabi.HtcnError, match="stale"):
        model.loss(sc1)                                  # its hout workspace now holds the second batch
    with pytest.raises(cabi.HtcnError, match="stale"):
        sc1.materialize()
    assert abs(float(model.loss(sc2)["scalars"][0].item()) - l1) <= 1e-5 * abs(l1)      # user order does not matter
    shard, n0, n1 = make_sharded_model(a, w, 1, 2, "f32")
    sc3, _ = shard.forward(x, y, m, s0)
    with pytest.raises(cabi.HtcnError, match="shard"):
        shard.loss(sc3)
