"""Hyper-parameter surface of the hot path: the reference's ``args.py`` flag names and defaults
(reference args.py:3-193), typed this time (the reference's flags mostly lack ``type=`` and arrive
as strings from the CLI -- SURVEY A.8 quirk 15).

Only the flags the HierTCN hot path reads are kept; reference-only paths (Pinterest data dirs,
html, caches) are out of scope.  Unlike the reference, nothing is parsed at import time:
call ``make_args(argv)``; ``make_args([])`` gives the reference defaults with
``model_type='hier', model_low_type='tcn'`` (the configuration the hot path is).
"""
from __future__ import annotations

from argparse import ArgumentParser, Namespace


def _int_list(s):
    if isinstance(s, (list, tuple)):
        return [int(v) for v in s]
    return [int(v) for v in str(s).replace("[", "").replace("]", "").split(",") if v.strip()]


def make_parser() -> ArgumentParser:
    p = ArgumentParser("hiertcn_b200")
    p.add_argument("--dataset", dest="dataset", default="xing")                         # args.py:6
    p.add_argument("--warm_start", dest="warm_start", action="store_true")              # :10
    # training
    p.add_argument("--lr", dest="learning_rate", default=1e-2, type=float)              # :13
    p.add_argument("--lr_schedule_no", dest="lr_schedule", action="store_false")        # :15
    p.add_argument("--epoch_max", dest="epoch_max", default=200, type=int)              # :17
    p.add_argument("--batch_size", dest="batch_size", default=32, type=int)             # :19
    p.add_argument("--has_batchnorm", dest="has_batchnorm", action="store_true")        # :22
    p.add_argument("--has_weightnorm", dest="has_weightnorm", action="store_true")      # :24
    p.add_argument("--has_layernorm", dest="has_layernorm", action="store_true")        # :26
    # network -- the reference defaults to the mv_xing baseline / gru low level (args.py:29,40);
    # this package IS the hier+tcn path, so those are the defaults here.
    p.add_argument("--model_type", dest="model_type", default="hier")
    p.add_argument("--model_low_type", dest="model_low_type", default="tcn")
    p.add_argument("--item_num", dest="item_num", default=20777 + 1, type=int)          # :47
    p.add_argument("--hidden_dim", dest="hidden_dim", default=128, type=int)            # :51
    p.add_argument("--num_layer", dest="num_layer", default=2, type=int)                # :53
    p.add_argument("--tcn_channel", dest="tcn_channel", default=[128, 128], type=_int_list)  # :56
    p.add_argument("--kernel_size", dest="kernel_size", default=5, type=int)            # :58
    p.add_argument("--strides", dest="strides", default=1, type=int)                    # :60
    p.add_argument("--dropout", dest="dropout", default=0.0, type=float)                # :64
    p.add_argument("--has_impression", dest="has_impression", action="store_true")      # :88
    p.add_argument("--has_gap", dest="has_gap", action="store_true")                    # :90
    p.add_argument("--gap_bandwidth", dest="gap_bandwidth", default=168, type=float)    # :92
    p.add_argument("--train_gap", dest="train_gap", action="store_true")                # :94
    p.add_argument("--data_noise", dest="data_noise", default=None)                     # :96
    p.add_argument("--input_dim", dest="input_dim", default=512, type=int)              # :100
    p.add_argument("--output_dim", dest="output_dim", default=None, type=int)           # :102 (= item_num)
    # loss
    p.add_argument("--loss", dest="loss", default="cross_entropy")                      # :106
    p.add_argument("--rank_metric", dest="rank_metric", default="l2")                   # :110
    p.add_argument("--num_neg_sample", dest="num_neg_sample", default=20, type=int)     # :113
    p.add_argument("--nce_weight", dest="nce_weight", default=1.0, type=float)          # :115
    p.add_argument("--l2_normalize", dest="l2_normalize", action="store_true")          # :117
    p.add_argument("--hinge_delta", dest="hinge_delta", default=0.1, type=float)        # :119
    # data
    p.add_argument("--max_seq_len", dest="max_seq_len", default=500, type=int)          # :123
    p.add_argument("--max_impression_len", dest="max_impression_len", default=20, type=int)
    p.add_argument("--max_activity_len", dest="max_activity_len", default=20, type=int)  # :127
    p.add_argument("--max_session_num", dest="max_session_num", default=10, type=int)   # :129
    p.add_argument("--epoch_batches_train", dest="epoch_batches_train", default=500, type=int)
    p.add_argument("--save_epoch", dest="save_epoch", default=20, type=int)
    p.add_argument("--test_epoch", dest="test_epoch", default=20, type=int)
    p.add_argument("--load_epoch", dest="load_epoch", default=20, type=int)
    p.add_argument("--shuffle", dest="shuffle", action="store_true")
    p.add_argument("--shuffle_no", dest="shuffle", action="store_false")
    # --- additions of this implementation (not in the reference) ---
    p.add_argument("--emb_dim", dest="emb_dim", default=128, type=int,
                   help="item-embedding width; hard-coded 128 in the reference (model_hier.py:50,85)")
    p.add_argument("--precision", dest="precision", default="bf16", choices=["f32", "bf16"],
                   help="arithmetic tier of the CUDA path: f32 (FFMA, 1e-4) or bf16 (tcgen05, 2e-2)")
    p.add_argument("--topk", dest="topk", default=100, type=int, help="k of the fused catalog top-k")
    p.set_defaults(shuffle=False)
    return p


def args_adjust(args: Namespace) -> Namespace:
    """reference args.py:293-316, restricted to what the hot path reads."""
    if args.model_type == "tcn":                      # single-level TCN (args.py:310-311)
        args.tcn_channel = [128, 128, 128, 128, 256, 256]
    if args.output_dim is None:
        args.output_dim = args.item_num
    return args


def make_args(argv=None) -> Namespace:
    """Parse ``argv`` (default: no flags -- NOT sys.argv, so importing never touches the CLI)."""
    args = make_parser().parse_args([] if argv is None else list(argv))
    return args_adjust(args)
