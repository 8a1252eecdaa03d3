"""ctypes binding of libhtcn.so (include/htcn.h).  No torch types cross this boundary: callers pass
raw device pointers (``tensor.data_ptr()``), sizes and a stream handle.

There is no fallback: if the library is missing ``load()`` raises, and if a call fails the status is
turned into ``HtcnError`` with the library's message."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libhtcn.so")

HTCN_F32, HTCN_BF16, HTCN_F32_W256 = 0, 1, 2
SCORE_CE, SCORE_RANK, SCORE_TOPK = 1, 2, 4
LOSS_KINDS = {"nce": 0, "hinge_sigmoid": 1, "hinge_logsigmoid": 2, "hinge_linear": 3, "bpr": 4}
MAX_TOPK = 128
WT_PITCH_BF16 = 144
K2TC_SCRATCH_BYTES = 8 * 2 * 128 * 128 * 2          # HTCN_K2TC_SCRATCH_BYTES


def tcn_scratch_floats(n_levels, K):
    """HTCN_TCN_SCRATCH_BYTES / 4: bf16 weight tiles (K taps + a possible down-sample kernel per level) + tables"""
    return ((1 + n_levels * (K + 1)) * 128 * 128 * 2 + 640 + 2 * 8 * 512 + 256 + 592 * 2 * n_levels * 8192 + 3) // 4


def gru_scratch_bytes(B):
    return 14 * 128 * 128 * 2 + 4096          # HTCN_GRU_SCRATCH_BYTES: independent of B (the fp32 state lives on chip)

_p, _i, _u, _f = C.c_void_p, C.c_int32, C.c_uint32, C.c_float
_pp = C.POINTER(C.c_void_p)        # host array of device pointers
_ip = C.POINTER(C.c_int32)         # host int array
_lp = C.POINTER(C.c_int64)         # host int64 array

# name -> argtypes; every function returns int32 status unless noted.  Keep in sync with include/htcn.h
# (tests/test_cabi.py parses the header and compares).
SIGNATURES = {
    "htcn_gather_meanpool": [_p, _i, _p, _i, _p, _p, _ip, _i, _i, _i, _p, _i, _p, _p],
    "htcn_gru_sessions": [_p, _p, _p, _pp, _pp, _pp, _pp, _i, _p, _i, _i, _i, _p, _p, _p, _p, _p],
    "htcn_tcn_forward": [_p, _i, _i, _p, _p, _pp, _pp, _pp, _pp, _i, _i, _ip, _i, _i, _i, _p, _p, _i, _p, _p],
    "htcn_tcn_forward_wide": [_p, _i, _p, _p, _pp, _pp, _pp, _pp, _ip, _i, _i, _ip, _i, _i, _i, _p, _p, C.c_int64, _p, _p],
    "htcn_prepare_wout": [_p, _p, _i, _p, _i, _p],
    "htcn_score_ce_rank_topk": [_p, _i, _i, _p, _p, _i, _i, _p, _p, _i, _u, _i, _i, _p, _p, _p, _p, _p, _p],
    "htcn_score_ce_rank_folded": [_p, _i, _p, _i, _i, _p, _p, _u, _i, _p, _p, _p, _p, _p],
    "htcn_score_logits": [_p, _i, _i, _p, _i, _p, _i, _p, _p],
    "htcn_target_logit": [_p, _i, _i, _p, _p, _i, _i, _p, _p, _p],
    "htcn_score_finish": [_p, _p, _p, _i, _i, _p, _p, _p, _p, _p],
    "htcn_topk_merge": [_p, _p, _i, _i, _i, _p, _p, _p],
    "htcn_score_ce_repair": [_p, _i, _i, _p, _i, _p, _p, _p, _p],
    "htcn_score_ce_repair_shard": [_p, _i, _i, _p, _i, _p, _p, _i, _p, _p],
    "htcn_score_topk": [_p, _i, _i, _p, _p, _i, _i, _i, _i, _p, C.c_int64, _p, _p, _p, _p],
    "htcn_score_ce_rank_topk_fused": [_p, _i, _i, _p, _p, _i, _i, _p, _p, _i, _i, _p, C.c_int64, _p, _p, _p, _p, _p, _p, _p],
    "htcn_catalog_gram": [_p, _i, _p, _i, _p, _p, _p],
    "htcn_logit_rownorm": [_p, _i, _i, _p, _f, _p, _p],
    "htcn_score_ce_rank_l2norm": [_p, _i, _i, _p, _p, _i, _i, _p, _p, _p, _i, _p, _p, _p, _p, _p],
    "htcn_scale_rows": [_p, _p, C.c_int64, _i, _p],
    "htcn_loss_metrics_reduce": [_p, _p, _p, _p, _i, _i, _i, _p, _p, _p, _p, _p, _p],
    "htcn_sampled_rank_loss": [_p, _i, _i, _p, _p, _p, _i, _i, _f, _f, _i, _p, _p],
    "htcn_sampled_rank_loss_wt": [_p, _i, _p, _p, _p, _i, _i, _f, _f, _i, _p, _p],
    "htcn_sampled_rank_loss_backward": [_p, _i, _i, _p, _p, _p, _i, _i, _f, _f, _i, _p, _p, _p, _p],
    "htcn_calc_score": [_p, _i, _i, _p, _p, _i, _i, _p, _p],
    # training step
    "htcn_loss_row_weights": [_p, _p, _i, _i, _p, _p],
    "htcn_score_ce_backward": [_p, _i, _i, _p, _p, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p],
    "htcn_score_ce_backward_bf16": [_p, _p, C.c_int64, _i, _p, _p, C.c_int64, _p, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p],
    "htcn_score_ce_fwd_bwd_bf16": [_p, _p, C.c_int64, _i, _p, _p, C.c_int64, _p, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p],
    "htcn_cast_transpose_bf16": [_p, _i, C.c_int64, _p, _p, C.c_int64, _p],
    "htcn_tcn_forward_train_bf16": [_p, _p, _p, _pp, _pp, _pp, _pp, _i, _i, _ip, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p],
    "htcn_tcn_forward_train": [_p, _p, _p, _pp, _pp, _pp, _pp, _i, _i, _ip, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p],
    "htcn_tcn_backward": [_p, _p, _p, _i, _p, _pp, _pp, _i, _i, _ip, _i, _i, _i, _p, _p, _p, _p, _p, _pp, _pp, _pp, _pp, _p, _p, _p, _p],
    "htcn_gru_sessions_train": [_p, _p, _p, _pp, _pp, _pp, _pp, _i, _p, _i, _i, _p, _p, _p, _p, _p],
    "htcn_gru_sessions_train_bf16": [_p, _p, _p, _pp, _pp, _pp, _pp, _i, _p, _i, _i, _p, _p, _p, _p, _p, _p],
    "htcn_gru_backward": [_p, _p, _p, _p, _pp, _pp, _i, _p, _i, _i, _p, _p, _pp, _pp, _pp, _pp, _p, _p, _p],
    "htcn_gather_backward": [_p, _p, _p, _p, _ip, _i, _i, _i, _i, _p, _p, _p],
    "htcn_adam_step": [_p, _p, _p, _p, C.c_int64, _f, _f, _f, _f, _p, _i, _p],
    "htcn_refresh_wout": [_p, _p, _i, _p, _i, _p],
    "htcn_assemble_batch": [_p, _p, _p, _p, _i, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _p, _p],
    # peer-memory exchanges of the catalog-sharded path
    "htcn_peer_alloc": [C.c_int64, _pp],
    "htcn_peer_free": [_p],
    "htcn_peer_export": [_p, _p],
    "htcn_peer_import": [_p, _pp],
    "htcn_peer_unimport": [_p],
    "htcn_peer_exchange": [_pp, _pp, _lp, _ip, _lp, _lp, _i, _i, _i, _pp, _p, _p, _p, _u, _p],
    "htcn_peer_bcast_owned": [_p, _p, _i, _i, _i, _pp, _i, _pp, _p, _p, _p, _u, _p],
    "htcn_peer_allreduce_adam": [_pp, _pp, _pp, _i, _i, _p, _p, _p, C.c_int64, _f, _f, _f, _f, _p, _pp, _pp, _p, _p, _p, _p, _u, _p],
}
PLAIN = {"htcn_abi_version": (C.c_int32, []), "htcn_last_error": (C.c_char_p, []),
         "htcn_device_ok": (C.c_int32, []),
         "htcn_topk_workspace_bytes": (C.c_int64, [_i, _i, _i, _i, _i]),
         "htcn_gru_backward_scratch_floats": (C.c_int64, [_i, _i, _i]),
         "htcn_batcher_scratch_ints": (C.c_int64, [_i, _i]),
         "htcn_catalog_gram_scratch_floats": (C.c_int64, []),
         "htcn_tcn_backward_tc_scratch_bytes": (C.c_int64, [_i, _i, _i, _i, _i])}


class HtcnError(RuntimeError):
    pass


_lib = None


def load(path: str | None = None):
    """dlopen libhtcn.so and declare every prototype.  Raises if the library is absent or a symbol is
    missing -- the product path must fail loudly rather than fall back."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise HtcnError("libhtcn.so not found at %s -- build it with `python -m hiertcn_b200.build` "
                        "(there is no CPU fallback)" % path)
    lib = C.CDLL(path)
    for name, (res, args) in PLAIN.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = C.c_int32, args
    _lib = lib
    return lib


# kernels launched per successful call (bench.py's gpu_launches claim); htcn_tcn_forward launches
# n_levels + 1 (counted by the caller through note_launches)
LAUNCHES_PER_CALL = {"htcn_gather_meanpool": 2, "htcn_gru_sessions": 1, "htcn_tcn_forward": 0,
                     "htcn_prepare_wout": 1, "htcn_score_ce_rank_topk": 2, "htcn_score_ce_rank_folded": 3, "htcn_score_logits": 1,
                     "htcn_target_logit": 1, "htcn_score_finish": 1, "htcn_topk_merge": 1, "htcn_score_topk": 5, "htcn_score_ce_repair": 1,
                     "htcn_score_ce_repair_shard": 1, "htcn_score_ce_rank_topk_fused": 6,
                     "htcn_catalog_gram": 2, "htcn_logit_rownorm": 1, "htcn_score_ce_rank_l2norm": 2, "htcn_scale_rows": 1,
                     "htcn_loss_metrics_reduce": 2, "htcn_sampled_rank_loss": 1, "htcn_sampled_rank_loss_wt": 1, "htcn_calc_score": 1,
                     "htcn_sampled_rank_loss_backward": 1,
                     # training step (the per-call counts of the multi-launch entry points are added by the caller)
                     "htcn_loss_row_weights": 1, "htcn_score_ce_backward": 1, "htcn_gru_sessions_train": 1, "htcn_gru_sessions_train_bf16": 2,
                     "htcn_gather_backward": 2, "htcn_adam_step": 1, "htcn_refresh_wout": 1,
                     "htcn_score_ce_backward_bf16": 3, "htcn_score_ce_fwd_bwd_bf16": 6, "htcn_cast_transpose_bf16": 1, "htcn_assemble_batch": 4,
                     "htcn_peer_exchange": 1, "htcn_peer_bcast_owned": 1, "htcn_peer_allreduce_adam": 1}
launch_count = 0


def note_launches(n: int):
    global launch_count
    launch_count += n


def call(name: str, *args):
    global launch_count
    lib = load()
    rc = getattr(lib, name)(*args)
    launch_count += LAUNCHES_PER_CALL.get(name, 0)
    if rc != 0:
        raise HtcnError("%s failed (%d): %s" % (name, rc, lib.htcn_last_error().decode()))


def score_fold_ws_bytes(Q):
    return (Q * 128 * 2 + 255) // 256 * 256 + Q * 4               # HTCN_SCORE_FOLD_WS_BYTES


def ce_bwd_bf16_ws_floats(Q, N):
    return 4 * (-(-Q // 64) * 64) + (-(-N // 64) * 64)            # HTCN_CE_BWD_BF16_WS_FLOATS


def ptr_array(ptrs):
    """host array of device pointers"""
    arr = (C.c_void_p * len(ptrs))(*[C.c_void_p(int(p)) for p in ptrs])
    return C.cast(arr, _pp), arr       # keep `arr` alive while the call runs


def long_array(vals):
    arr = (C.c_int64 * len(vals))(*[int(v) for v in vals])
    return C.cast(arr, _lp), arr


def int_array(vals):
    arr = (C.c_int32 * len(vals))(*[int(v) for v in vals])
    return C.cast(arr, _ip), arr
