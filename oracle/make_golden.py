"""Generate tests/golden/*.npz by running the REFERENCE's own python under oracle/tf_shim.

TEST INFRASTRUCTURE.  Runs only where /root/reference exists (this dev container); the fixtures it
writes are committed and are what travels to the GPU box.

    python oracle/make_golden.py            # regenerate every fixture
    python oracle/make_golden.py --check    # regenerate in memory and compare with the committed files

What is executed unmodified from the reference: args.py, loss.py (calc_loss, calc_score,
calc_metric, calc_metric_fast), model_hier.py (model_hier), model_tcn.py (model_tcn),
customized_tcn_cell.py (CausalConv1D, TemporalBlock, TemporalConvNet).  What is substituted:
``tensorflow`` (oracle/tf_shim/tensorflow: numpy restatement of the ops used) and the vendored-TF
layer files customized_{dense,convolution}_layer.py / customed_gru_cell.py plus non-hot-path
modules (oracle/tf_shim/stubs).  model.py (graph assembly at import, placeholders) cannot be run;
its mask/loss reduction (model.py:98-117) is restated in oracle/hiertcn_oracle.py:hier_loss.
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("HTCN_REFERENCE", "/root/reference")
GOLD = os.path.join(ROOT, "tests", "golden")


def import_reference():
    """Put the shim + stubs ahead of the reference dir and import the reference modules."""
    if not os.path.isdir(REF):
        raise SystemExit("reference checkout not found at %s" % REF)
    for p in (REF, os.path.join(HERE, "tf_shim", "stubs"), os.path.join(HERE, "tf_shim")):
        sys.path.insert(0, p)
    argv, sys.argv = sys.argv, ["run_xing.py", "--model_type", "hier", "--model_low_type", "tcn"]
    try:
        import tensorflow as tf          # the shim
        import args as ref_args          # reference args.py (parses sys.argv at import, args.py:318)
        import loss as ref_loss          # reference loss.py
        import model_hier as ref_hier    # reference model_hier.py (+ model_tcn.py, customized_tcn_cell.py)
    finally:
        sys.argv = argv
    assert tf.__file__.startswith(HERE), "real tensorflow shadowed the shim?"
    for m in (ref_args, ref_loss, ref_hier):
        assert m.__file__.startswith(REF), m.__file__
    return tf, ref_args.args, ref_loss, ref_hier


def _set_args(a, **kw):
    for k, v in kw.items():
        setattr(a, k, v)


def hier_case(tf, a, ref_loss, ref_hier, *, name, N, B, S, Lmax, hidden_dim, num_layer, tcn_channel,
              kernel_size, seed, store_weights, kernel_scale=2.0, has_gap=False, l2_normalize=False, warm_start=False):
    """``has_gap``: the gap decay of model_hier.py:40-47 (train_gap off) with random gaps; ``l2_normalize``: the normalised
    head of model_tcn.py:42-43; ``warm_start``: a random mask_warmstart multiplied into mask_y (model.py:102-103)."""
    sys.path.insert(0, ROOT)
    from hiertcn_b200.data_loader import synthetic_batch
    from hiertcn_b200.weights import hier_weight_shapes, init_weights, weights_sha256

    _set_args(a, item_num=N, output_dim=N, hidden_dim=hidden_dim, num_layer=num_layer,
              tcn_channel=list(tcn_channel), kernel_size=kernel_size, dropout=0.0, loss="cross_entropy",
              model_type="hier", model_low_type="tcn", has_gap=has_gap, train_gap=False, gap_bandwidth=168.0,
              l2_normalize=l2_normalize, warm_start=warm_start)
    shapes = hier_weight_shapes(N, hidden_dim, num_layer, tcn_channel, kernel_size)
    w = init_weights(shapes, seed=seed, kernel_scale=kernel_scale, bias_noise=0.1)
    x_list, y_list, mask_list = synthetic_batch(B, S, Lmax, N, seed=seed + 1, lengths="ragged",
                                                id_dist="uniform", mask_keep=0.6)
    rng = np.random.default_rng(seed + 2)
    state0 = rng.normal(0, 0.5, size=(B, hidden_dim * num_layer)).astype(np.float32)
    x_gap = [rng.exponential(120.0, size=(B, 1)).astype(np.float32) for _ in range(S)] if has_gap else None
    T_all = sum(x.shape[1] for x in x_list)
    mask_warm = (rng.random((B, T_all)) < 0.7).astype(np.float32) if warm_start else None

    # ---- run the reference graph code eagerly (float64 so the fixture is a precise target) ----
    res = {}
    for dt, tag in ((np.float64, "f64"), (np.float32, "f32")):
        tf.WEIGHTS = {k: v.astype(dt) for k, v in w.items()}
        del tf.TOUCHED[:]
        x_ids = [x.astype(np.int32) for x in x_list]
        y_ids = [y.astype(np.int32) for y in y_list]
        # model.py:59-61
        xs = [tf.one_hot(t, depth=N, dtype=dt) * tf.cast(tf.expand_dims(tf.sign(t), axis=-1), dtype=dt) for t in x_ids]
        ys = [tf.one_hot(t, depth=N, dtype=dt) * tf.cast(tf.expand_dims(tf.sign(t), axis=-1), dtype=dt) for t in y_ids]
        masks = [m.astype(dt) for m in mask_list]
        gaps = [g.astype(dt) for g in x_gap] if has_gap else None
        pred, state = ref_hier.model_hier(a, xs, ys, masks, state0.astype(dt), x_gap=gaps, training=np.asarray(False))
        y_id = np.concatenate(y_ids, 1)
        y = tf.one_hot(y_id, depth=N, dtype=dt) * tf.cast(tf.expand_dims(tf.sign(y_id), axis=-1), dtype=dt)
        # model.py:62,105-117 (restated -- model.py itself cannot be imported)
        mask_y = tf.cast(tf.sign(y_id), dtype=dt)
        if warm_start:
            mask_y = mask_y * mask_warm.astype(dt)               # model.py:102-103
        pred = pred * tf.expand_dims(mask_y, -1)
        loss_bt = ref_loss.calc_loss(pred, y)                    # reference loss.py:20-21
        loss_bt = loss_bt * mask_y
        activity_count = tf.reduce_sum(mask_y, 1)
        user_count = tf.reduce_sum(tf.sign(activity_count), axis=-1)
        activity_count = activity_count + dt(1e-6)
        loss = tf.reduce_sum(tf.reduce_sum(loss_bt, 1) / activity_count) / user_count
        met = ref_loss.calc_metric_fast(pred, mask_y, activity_count, user_count, y)   # loss.py:163-221
        res[tag] = dict(pred=pred, state=state, loss=loss, loss_bt=loss_bt, rec1=met[0], rec5=met[1],
                        rec10=met[2], mrr=met[3], mrp=met[4], ranks_float=met[5], ranks=met[6])
        touched = list(tf.TOUCHED)
    assert sorted(touched) == sorted(shapes.keys()), (
        "variables the reference graph touched differ from the weight contract:\n%s\nvs\n%s"
        % (sorted(touched), sorted(shapes.keys())))

    out = dict(
        N=N, B=B, S=S, hidden_dim=hidden_dim, num_layer=num_layer, tcn_channel=np.asarray(tcn_channel),
        kernel_size=kernel_size, weight_seed=seed, kernel_scale=kernel_scale, bias_noise=0.1,
        weights_sha256=weights_sha256(w), state0=state0,
        var_names=np.asarray(touched),
    )
    for s in range(S):
        out[f"x_{s}"] = x_list[s].astype(np.int32)
        out[f"y_{s}"] = y_list[s].astype(np.int32)
        out[f"mask_{s}"] = mask_list[s].astype(np.float32)
    r64, r32 = res["f64"], res["f32"]
    out.update(pred_f64=r64["pred"].astype(np.float32), state_f64=r64["state"], loss_f64=r64["loss"],
               loss_bt_f64=r64["loss_bt"], ranks_f64=r64["ranks"],
               metrics_f64=np.asarray([r64[k] for k in ("rec1", "rec5", "rec10", "mrr", "mrp")]),
               loss_f32=r32["loss"], state_f32=r32["state"], ranks_f32=r32["ranks"],
               metrics_f32=np.asarray([r32[k] for k in ("rec1", "rec5", "rec10", "mrr", "mrp")]))
    if has_gap:
        out.update(gap_bandwidth=168.0, **{f"x_gap_{s}": x_gap[s] for s in range(S)})
    if warm_start:
        out["mask_warmstart"] = mask_warm
    out["l2_normalize"] = int(l2_normalize)
    if store_weights:
        for k, v in w.items():
            out["w|" + k.replace("/", "|")] = v
    return name, out


def loss_case(tf, a, ref_loss, seed=7):
    """Known-answer vectors for every branch of reference loss.py that the hot path names."""
    rng = np.random.default_rng(seed)
    B, T, d, k, N = 3, 5, 16, 20, 37
    pred = rng.normal(size=(B, T, d)).astype(np.float64)
    y = rng.normal(size=(B, T, d)).astype(np.float64)
    y_imp = rng.normal(size=(B, T, k, d)).astype(np.float64)
    out = dict(pred=pred, y=y, y_impression=y_imp, num_neg_sample=k, hinge_delta=0.1, nce_weight=1)
    _set_args(a, num_neg_sample=k, hinge_delta=0.1, nce_weight=1, max_impression_len=k)
    for name in ("l2", "nce", "hinge_sigmoid", "hinge_logsigmoid", "hinge_linear", "bpr"):
        _set_args(a, loss=name)
        out["loss_" + name] = np.asarray(ref_loss.calc_loss(pred, y, y_imp))        # loss.py:18-71
    for rm in ("l2", "inner_prod"):
        _set_args(a, rank_metric=rm)
        out["score_" + rm] = np.asarray(ref_loss.calc_score(pred, y_imp))            # loss.py:76-105
    # calc_metric (full top_k ordering, loss.py:108-160) and calc_metric_fast on catalog-style scores,
    # with deliberate ties to pin the tie rule of the restated top_k
    score = np.round(rng.normal(size=(B, T, N)) * 4) / 4
    y_id = rng.integers(0, N, size=(B, T)).astype(np.int32)
    y_id[0, -2:] = 0
    mask_y = np.sign(y_id).astype(np.float64)
    score = score * mask_y[..., None]
    act = mask_y.sum(1)
    ucount = np.sign(act).sum()
    act = act + 1e-6
    _set_args(a, item_num=N, model_type="hier", loss="cross_entropy")
    y_oh = tf.one_hot(y_id, depth=N, dtype=np.float64) * np.sign(y_id)[..., None]
    mf = ref_loss.calc_metric_fast(score, mask_y, act, ucount, y_oh)
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        m = ref_loss.calc_metric(score, mask_y, act, ucount, y_id)
    out.update(score=score, y_id=y_id, metric_fast_scalars=np.asarray(mf[:5]), metric_fast_ranks_float=mf[5],
               metric_fast_ranks=mf[6], metric_scalars=np.asarray(m[:5]), metric_ranks_float=m[5],
               metric_topk_indices=np.asarray(m[6]), metric_topk_values=np.asarray(m[7]), metric_ranks=m[8])
    return "loss_vectors", out


def build_all():
    tf, a, ref_loss, ref_hier = import_reference()
    cases = []
    cases.append(hier_case(tf, a, ref_loss, ref_hier, name="hier_default_arch", N=61, B=3, S=3, Lmax=6,
                           hidden_dim=128, num_layer=2, tcn_channel=(128, 128), kernel_size=5, seed=11,
                           store_weights=True))
    cases.append(hier_case(tf, a, ref_loss, ref_hier, name="hier_downsample_3lvl", N=40, B=4, S=4, Lmax=9,
                           hidden_dim=16, num_layer=2, tcn_channel=(32, 32, 48), kernel_size=3, seed=23,
                           store_weights=True))
    # optional variants of SURVEY 8(f-4): gap decay, l2-normalised head, warm-start loss mask
    cases.append(hier_case(tf, a, ref_loss, ref_hier, name="hier_gap_l2norm_warmstart", N=53, B=5, S=3, Lmax=7,
                           hidden_dim=128, num_layer=2, tcn_channel=(128, 128), kernel_size=5, seed=37,
                           store_weights=True, has_gap=True, l2_normalize=True, warm_start=True))
    cases.append(loss_case(tf, a, ref_loss))
    return cases


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    opt = ap.parse_args()
    os.makedirs(GOLD, exist_ok=True)
    for name, out in build_all():
        path = os.path.join(GOLD, name + ".npz")
        if opt.check:
            old = np.load(path)
            assert set(old.files) == set(out), (name, set(old.files) ^ set(out))
            for k, v in out.items():
                if np.asarray(v).dtype.kind in "fc":
                    np.testing.assert_allclose(old[k], v, rtol=1e-12, atol=0, err_msg=f"{name}:{k}")
                else:
                    assert (old[k] == np.asarray(v)).all(), f"{name}:{k}"
            print("ok", name)
        else:
            np.savez_compressed(path, **out)
            print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
