"""Weight contract of the hot path: a state dict keyed by the TF variable names the reference graph
creates (derived from scope usage in model_hier.py:28,50,85, model_tcn.py:26,35,41,
customized_tcn_cell.py:80-87,153-154, customed_gru_cell.py:315,328,1056,1170-1185), so that
"identical weights" is well defined against a real TF-1.x run.  numpy only (host side)."""
from __future__ import annotations

import hashlib
from collections import OrderedDict

import numpy as np


def hier_weight_shapes(item_num: int, hidden_dim: int = 128, num_layer: int = 2,
                       tcn_channel=(128, 128), kernel_size: int = 5, emb_dim: int = 128,
                       output_dim: int | None = None) -> "OrderedDict[str, tuple]":
    out_dim = item_num if output_dim is None else output_dim
    s: "OrderedDict[str, tuple]" = OrderedDict()
    s["hier/emb/kernel"] = (item_num, emb_dim)
    s["hier/emb/bias"] = (emb_dim,)
    s["hier/tcn/emb/kernel"] = (emb_dim + num_layer * hidden_dim, 128)      # model_tcn.py:35 units=128
    cin = 128
    for lvl, c in enumerate(tcn_channel):
        p = f"hier/tcn/temporal_conv_net/tblock_{lvl}"
        s[p + "/conv1/kernel"] = (kernel_size, cin, c)
        s[p + "/conv1/bias"] = (c,)
        if cin != c:                                                        # customized_tcn_cell.py:102-106
            s[p + "/dense/kernel"] = (cin, c)
            s[p + "/dense/bias"] = (c,)
        cin = c
    s["hier/tcn/dense/kernel"] = (cin, out_dim)
    s["hier/tcn/dense/bias"] = (out_dim,)
    inp = emb_dim
    for g in range(num_layer):
        p = f"hier/multi_rnn_cell/cell_{g}/gru_cell"
        s[p + "/gates/kernel"] = (inp + hidden_dim, 2 * hidden_dim)
        s[p + "/gates/bias"] = (2 * hidden_dim,)
        s[p + "/candidate/kernel"] = (inp + hidden_dim, hidden_dim)
        s[p + "/candidate/bias"] = (hidden_dim,)
        inp = hidden_dim
    return s


def tcn_weight_shapes(in_dim: int, tcn_channel=(128, 128, 128, 128), kernel_size: int = 5,
                      output_dim: int | None = None, scope: str = "tcn") -> "OrderedDict[str, tuple]":
    """Single-level model_tcn (model_tcn.py:26-44), BASELINE config 3."""
    s: "OrderedDict[str, tuple]" = OrderedDict()
    s[f"{scope}/emb/kernel"] = (in_dim, 128)
    cin = 128
    for lvl, c in enumerate(tcn_channel):
        p = f"{scope}/temporal_conv_net/tblock_{lvl}"
        s[p + "/conv1/kernel"] = (kernel_size, cin, c)
        s[p + "/conv1/bias"] = (c,)
        if cin != c:
            s[p + "/dense/kernel"] = (cin, c)
            s[p + "/dense/bias"] = (c,)
        cin = c
    if output_dim is not None:
        s[f"{scope}/dense/kernel"] = (cin, output_dim)
        s[f"{scope}/dense/bias"] = (output_dim,)
    return s


def _glorot_uniform(rng, shape):
    """[TF-sem] default initializer of tf.layers / get_variable: glorot_uniform."""
    if len(shape) == 2:
        fan_in, fan_out = shape
    else:                                   # conv kernel [K, Cin, Cout]
        rf = int(np.prod(shape[:-2]))
        fan_in, fan_out = shape[-2] * rf, shape[-1] * rf
    lim = np.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-lim, lim, size=shape).astype(np.float32)


def init_weights(shapes, seed: int = 1234, kernel_scale: float = 1.0, bias_noise: float = 0.0):
    """glorot-uniform kernels, zero biases, GRU gate bias 1.0 (customed_gru_cell.py:312-321).

    ``kernel_scale`` > 1 gives the "trained-like" variant (logits not near-uniform);
    ``bias_noise`` > 0 perturbs biases so tests cannot pass with a dropped bias."""
    rng = np.random.default_rng(seed)
    w = OrderedDict()
    for name, shape in shapes.items():
        if name.endswith("/kernel"):
            w[name] = _glorot_uniform(rng, shape) * np.float32(kernel_scale)
        elif name.endswith("gates/bias"):
            w[name] = np.ones(shape, np.float32)
        else:
            w[name] = np.zeros(shape, np.float32)
        if bias_noise and name.endswith("/bias"):
            w[name] = w[name] + rng.normal(0, bias_noise, size=shape).astype(np.float32)
    return w


def weights_sha256(w) -> str:
    h = hashlib.sha256()
    for k in sorted(w):
        h.update(k.encode())
        h.update(np.ascontiguousarray(w[k], dtype=np.float32).tobytes())
    return h.hexdigest()


def save_npz(path, w):
    np.savez(path, **{k.replace("/", "|"): v for k, v in w.items()})


def load_npz(path):
    z = np.load(path)
    return OrderedDict((k.replace("|", "/"), z[k]) for k in z.files)


def fold_weightnorm(w):
    """Weight-norm is a pure re-parameterisation (customized_dense_layer.py:140-142,
    customized_convolution_layer.py:146-148): fold ``g`` into the kernel at load time."""
    out = OrderedDict()
    for k, v in w.items():
        if k.endswith("/g"):
            continue
        if k.endswith("/kernel") and (k[:-len("kernel")] + "g") in w:
            g = w[k[:-len("kernel")] + "g"]
            if v.ndim == 3:
                nrm = np.sqrt(np.maximum((v * v).sum((0, 1), keepdims=True), 1e-12))
                v = g.reshape(1, 1, -1) * v / nrm
            else:
                v = g * v / np.sqrt((v * v).sum(0, keepdims=True))
        out[k] = v.astype(np.float32)
    return out
