"""Shared test helpers: load golden fixtures, rebuild their weights, small synthetic cases."""
import os

import numpy as np

from hiertcn_b200.weights import hier_weight_shapes, init_weights, weights_sha256

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_hier_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    S = int(z["S"])
    x = [z[f"x_{s}"] for s in range(S)]
    y = [z[f"y_{s}"] for s in range(S)]
    m = [z[f"mask_{s}"] for s in range(S)]
    wkeys = [k for k in z.files if k.startswith("w|")]
    if wkeys:
        w = {k[2:].replace("|", "/"): z[k] for k in wkeys}
    else:
        shapes = hier_weight_shapes(int(z["N"]), int(z["hidden_dim"]), int(z["num_layer"]),
                                    tuple(int(c) for c in z["tcn_channel"]), int(z["kernel_size"]))
        w = init_weights(shapes, seed=int(z["weight_seed"]), kernel_scale=float(z["kernel_scale"]),
                         bias_noise=float(z["bias_noise"]))
    assert weights_sha256(w) == str(z["weights_sha256"]), \
        "weight RNG drifted from the fixture: regenerate with oracle/make_golden.py"
    return z, x, y, m, w


def golden_variant_kwargs(z):
    """oracle kwargs of the optional variants a fixture was generated with (gap decay, l2-normalised head, warm-start mask)"""
    kw = {}
    if "x_gap_0" in z.files:
        kw.update(x_gap=[z[f"x_gap_{s}"] for s in range(int(z["S"]))], gap_bandwidth=float(z["gap_bandwidth"]))
    if "l2_normalize" in z.files and int(z["l2_normalize"]):
        kw["l2_norm"] = True
    if "mask_warmstart" in z.files:
        kw["mask_warmstart"] = z["mask_warmstart"]
    return kw


def small_case(B=5, S=3, L=7, N=97, seed=0, lengths="ragged", kernel_scale=2.0, tcn_channel=(128, 128),
               kernel_size=5, mask_keep=0.7):
    from hiertcn_b200.data_loader import synthetic_batch
    shapes = hier_weight_shapes(N, 128, 2, tcn_channel, kernel_size)
    w = init_weights(shapes, seed=100 + seed, kernel_scale=kernel_scale, bias_noise=0.1)
    x, y, m = synthetic_batch(B, S, L, N, seed=seed, lengths=lengths, id_dist="uniform", mask_keep=mask_keep)
    rng = np.random.default_rng(seed + 5)
    state0 = rng.normal(0, 0.5, size=(B, 256)).astype(np.float32)
    return x, y, m, state0, w
