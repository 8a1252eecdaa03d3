"""CPU tests of the host-side mirror: args surface, batch layout, queue loader, weight contract."""
import numpy as np

from hiertcn_b200.args import make_args
from hiertcn_b200.data_loader import (Dataloader_hier_model_xing, make_synthetic_interactions, pack_batch,
                                      synthetic_batch)
from hiertcn_b200.weights import fold_weightnorm, hier_weight_shapes, init_weights, load_npz, save_npz


def test_args_defaults_match_reference():
    a = make_args()
    # reference args.py defaults (file:line in hiertcn_b200/args.py)
    assert (a.batch_size, a.item_num, a.output_dim, a.hidden_dim, a.num_layer) == (32, 20778, 20778, 128, 2)
    assert a.tcn_channel == [128, 128] and a.kernel_size == 5 and a.strides == 1 and a.dropout == 0.0
    assert (a.max_activity_len, a.max_session_num, a.num_neg_sample, a.hinge_delta) == (20, 10, 20, 0.1)
    assert a.loss == "cross_entropy" and a.learning_rate == 1e-2 and not a.has_batchnorm
    b = make_args(["--tcn_channel", "128,128,128,128", "--batch_size", "4096", "--item_num", "1000000"])
    assert b.tcn_channel == [128] * 4 and b.batch_size == 4096 and b.output_dim == 1000000
    assert make_args(["--model_type", "tcn"]).tcn_channel == [128, 128, 128, 128, 256, 256]   # args.py:310-311


def test_synthetic_batch_layout():
    x, y, m = synthetic_batch(6, S=4, L=7, item_num=50, seed=3, lengths="ragged")
    assert len(x) == len(y) == len(m) == 4
    for xs, ys, ms in zip(x, y, m):
        assert xs.shape == ys.shape and ms.shape == (6, 1) and xs.dtype == np.float64
        assert (xs[:, 0] == 0).all()                               # null id first (data_loader.py:246-261)
        assert (xs[:, 1:] == ys[:, :-1]).all()                      # x is y shifted right by one
        n = (ys > 0).sum(1)
        assert (n >= 1).all() and n.max() == ys.shape[1]            # L_s = per-slot max
        for b in range(6):                                          # zero right-padding only
            assert (ys[b, :n[b]] > 0).all() and (ys[b, n[b]:] == 0).all()
        assert ys.max() < 50
    pk = pack_batch(x, y, m)
    assert pk["x_id"].dtype == np.int32 and pk["x_id"].shape == pk["y_id"].shape
    assert pk["slot_off"][0] == 0 and pk["slot_off"][-1] == pk["x_id"].shape[1] and pk["mask"].shape == (4, 6)


def test_queue_loader_semantics():
    table, data = make_synthetic_interactions(60, 80, seed=1)
    a = make_args(["--batch_size", "4", "--max_session_num", "3", "--max_activity_len", "5"])
    ld = Dataloader_hier_model_xing(a, "train", data=(table, data))
    seen_reset = False
    for _ in range(6):
        x, y, m, info = ld.get_batch()
        assert len(x) == 3 and all(v.shape[0] == 4 for v in x)
        for xs, ys, ms, inf in zip(x, y, m, info):
            assert xs.shape[1] <= 5 and inf.shape == ys.shape + (5,)
            assert (xs[:, 1:] == ys[:, :-1]).all()
            assert ((ys > 0).sum(1) >= 1).all()
            assert (inf[..., 1] == ys).all()                        # info carries the item id of each y
            seen_reset |= bool((ms == 0).any())
            # sessions of one slot never mix users: user_id constant over the valid positions
            for b in range(4):
                u = inf[b, ys[b] > 0, 0]
                assert (u == u[0]).all()
    assert seen_reset


def test_weight_contract_names_and_roundtrip(tmp_path):
    s = hier_weight_shapes(20778)
    assert sum(int(np.prod(v)) for v in s.values()) == 5750698       # SURVEY.md A.6 / BASELINE.md
    assert s["hier/tcn/emb/kernel"] == (384, 128) and s["hier/tcn/dense/kernel"] == (128, 20778)
    w = init_weights(hier_weight_shapes(30), seed=1)
    assert (w["hier/multi_rnn_cell/cell_0/gru_cell/gates/bias"] == 1).all()
    assert (w["hier/emb/bias"] == 0).all()
    save_npz(tmp_path / "w.npz", w)
    w2 = load_npz(tmp_path / "w.npz")
    assert list(w2) == list(w) and all((w2[k] == w[k]).all() for k in w)
    # weight-norm folds into the kernel
    w3 = dict(w)
    w3["hier/tcn/dense/g"] = np.full(30, 2.0, np.float32)
    f = fold_weightnorm(w3)
    np.testing.assert_allclose(np.sqrt((f["hier/tcn/dense/kernel"] ** 2).sum(0)), 2.0, rtol=1e-5)


def test_loader_state_dict_resumes_the_same_batches():
    import json
    table, data = make_synthetic_interactions(40, 101, seed=5)
    a = make_args(["--batch_size", "4", "--max_session_num", "3", "--max_activity_len", "5", "--shuffle"])
    ld = Dataloader_hier_model_xing(a, "train", data=(table, data))
    for _ in range(3):
        ld.get_batch()
    snap = json.loads(json.dumps(ld.state_dict()))           # survives a JSON round trip (checkpoint format)
    want = [ld.get_batch() for _ in range(4)]
    ld2 = Dataloader_hier_model_xing(a, "train", data=(table, data))
    ld2.load_state_dict(snap)
    got = [ld2.get_batch() for _ in range(4)]
    for (x1, y1, m1, i1), (x2, y2, m2, i2) in zip(want, got):
        for u, v in zip(x1 + y1 + m1 + i1, x2 + y2 + m2 + i2):
            np.testing.assert_array_equal(u, v)


def test_device_batcher_schedule_replays_the_loader_queues():
    """hiertcn_b200.device_batcher.build_schedule (host half of the GPU batcher) vs the queue loader, batch by batch"""
    from hiertcn_b200.device_batcher import build_schedule
    table, data = make_synthetic_interactions(60, 211, seed=8)
    for shuffle in (False, True):
        a = make_args(["--batch_size", "5", "--max_session_num", "3", "--max_activity_len", "6"] + (["--shuffle"] if shuffle else []))
        ld = Dataloader_hier_model_xing(a, "train", data=(table, data))
        n_train = int(table.shape[0] * 0.8)
        items, sess_off, sched, flag, lens = build_schedule(table[:n_train], data, 5, shuffle, getattr(a, "seed", 0), passes=2)
        B, S, L = 5, 3, 6
        for k in range(int(lens.min()) // S):
            x, y, m, _ = ld.get_batch()
            for s in range(S):
                for b in range(B):
                    sid = sched[b, k * S + s]
                    seq = items[sess_off[sid]:sess_off[sid + 1]][:L]
                    n = len(seq)
                    yy = np.zeros(L, np.int64)
                    xx = np.zeros(L, np.int64)
                    yy[:n] = seq
                    xx[1:min(n + 1, L)] = seq[:L - 1][:n]
                    w = y[s].shape[1]
                    np.testing.assert_array_equal(y[s][b], yy[:w])
                    np.testing.assert_array_equal(x[s][b], xx[:w])
                    assert (yy[w:] == 0).all()
                    assert m[s][b, 0] == 1 - flag[b, k * S + s]


def test_device_layout_padding_is_exact_and_round_trips():
    """hiertcn_b200.weights.to_device_layout zero-pads every width to 128: (1) it round-trips the TF names / shapes, and
    (2) the oracle evaluated on the PADDED model (as a 128-wide TF-named dict) reproduces the unpadded model exactly -- the
    padded channels / GRU units stay zero, so the kernels' 128-wide blocks compute the reference's function"""
    from helpers import load_hier_golden
    from hiertcn_b200.weights import from_device_layout, pad_state, to_device_layout, unpad_state
    from oracle import hiertcn_oracle as O
    for name in ("hier_default_arch", "hier_downsample_3lvl"):
        z, x, y, m, w = load_hier_golden(name)
        lay, meta = to_device_layout(w)
        back = from_device_layout(lay, meta)
        assert set(back) == set(w)
        for k in w:
            assert back[k].shape == w[k].shape and np.array_equal(back[k], w[k]), k
    assert meta["channels"] == [32, 32, 48] and meta["ds"] == [True, False, True] and meta["H"] == 16
    wide = from_device_layout(lay, dict(meta, ed=128, H=128, channels=[128] * 3))      # the padded model, TF-named
    ref = O.forward_loss_metrics(x, y, m, z["state0"], w, 2, "f64")
    pad = O.forward_loss_metrics(x, y, m, pad_state(z["state0"], meta), wide, 2, "f64")
    assert pad["loss"] == ref["loss"]
    assert np.array_equal(pad["pred"], ref["pred"]) and np.array_equal(unpad_state(pad["state"], meta), ref["state"])
    full = pad["state"].reshape(len(ref["state"]), 2, 128)
    assert (full[:, :, 16:] == 0).all(), "padded GRU units must stay exactly zero"


def test_device_layout_two_plane_levels_round_trip():
    """129..256-channel levels (the reference's single-level default [128,128,128,128,256,256], args.py:310-311) are laid
    out as 128-wide planes: conv kernels as [P_out, P_in*K, 128, 128] blocks; the layout round-trips the TF shapes"""
    from hiertcn_b200.args import make_args
    from hiertcn_b200.weights import from_device_layout, init_weights, tcn_weight_shapes, to_device_layout
    a = make_args(["--model_type", "tcn"])
    assert list(a.tcn_channel) == [128, 128, 128, 128, 256, 256]
    w = init_weights(tcn_weight_shapes(300, tuple(a.tcn_channel), 5, 300, "tcn"), seed=1, bias_noise=0.1)
    lay, meta = to_device_layout(w, "tcn")
    assert meta["planes"] == [1, 1, 1, 1, 2, 2] and meta["wide"] and meta["ds"] == [False] * 4 + [True, False]
    assert lay["conv_w4"].shape == (2, 5, 128, 128) and lay["conv_w5"].shape == (2, 10, 128, 128)
    assert lay["ds_w4"].shape == (2, 1, 128, 128) and lay["w_out"].shape == (256, 300)
    # block (po, pi*K + tap) = W[tap][pi*128.., po*128..]
    k5 = w["tcn/temporal_conv_net/tblock_5/conv1/kernel"]
    assert np.array_equal(lay["conv_w5"][1, 5 + 2], k5[2, 128:256, 128:256])
    back = from_device_layout(lay, meta)
    assert set(back) == set(w) and all(np.array_equal(back[k], w[k]) for k in w)
