"""Minimal numpy stand-in for the TensorFlow-1.6 ops the HierTCN hot path calls.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Purpose: let the reference's own
model_hier.py / model_tcn.py / customized_tcn_cell.py / loss.py run UNMODIFIED (eagerly, on
numpy arrays) so that their control flow -- which ops, in which order, on which tensors, under
which variable names -- pins the oracle.  The numerical meaning of each op below is restated from
TF-1.6 semantics [TF-sem]; that part remains unpinned.

Variables are not created: ``get_variable`` looks the fully scoped name up in
``tensorflow.WEIGHTS`` (a dict keyed by TF variable names), and raises KeyError on a
name the weight contract does not know -- which is itself a check of the scoping rules.
"""
import contextlib
import re
import sys
import types

import numpy as np

float32 = np.float32
float64 = np.float64
int32 = np.int32
int64 = np.int64
bool = np.bool_  # noqa: A001
AUTO_REUSE = "AUTO_REUSE"

WEIGHTS = {}          # name -> ndarray, set by the golden generator
TOUCHED = []          # variable names in first-use order (checked against the weight contract)
_SCOPE = []


class _T(np.ndarray):
    """ndarray that also answers get_shape() (used in reference print statements)."""

    def get_shape(self):
        return tuple(self.shape)


def _t(a):
    return np.asarray(a).view(_T)


@contextlib.contextmanager
def variable_scope(name, reuse=None, **_):
    _SCOPE.append(name)
    try:
        yield name
    finally:
        _SCOPE.pop()


name_scope = variable_scope


def current_scope():
    return "/".join(_SCOPE)


def get_variable(name, shape=None, dtype=None, initializer=None, trainable=True, **_):
    full = "/".join(_SCOPE + [name])
    if full not in WEIGHTS:
        raise KeyError("tf_shim: variable %r is not in the weight contract" % full)
    if full not in TOUCHED:
        TOUCHED.append(full)
    v = WEIGHTS[full]
    if shape is not None and tuple(int(s) for s in shape) != tuple(v.shape):
        raise ValueError("tf_shim: %s has shape %s, graph asks for %s" % (full, v.shape, tuple(shape)))
    return v


def constant(value, dtype=None, shape=None, name=None):
    return np.asarray(value, dtype=dtype)


def zeros_initializer(*a, **k):
    return "zeros"


def random_normal(shape, mean=0.0, stddev=1.0, dtype=np.float32, **_):
    raise NotImplementedError("data_noise is off on the hot path")


# ---- elementwise / shape ops ------------------------------------------------------------
def sign(x):
    return np.sign(x)


abs = np.abs  # noqa: A001
exp = np.exp
log = np.log
square = np.square
sqrt = np.sqrt


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def cast(x, dtype):
    return np.asarray(x).astype(dtype)


def shape(x):
    return np.asarray(np.shape(x))


def ones(shape, dtype=np.float32):
    return np.ones(tuple(int(s) for s in np.atleast_1d(shape)), dtype=dtype)


def zeros(shape, dtype=np.float32):
    return np.zeros(tuple(int(s) for s in np.atleast_1d(shape)), dtype=dtype)


def expand_dims(x, axis=None, dim=None):
    return np.expand_dims(x, axis if axis is not None else dim)


def squeeze(x, axis=None):
    return np.squeeze(x, axis=axis)


def tile(x, multiples):
    return np.tile(x, tuple(int(m) for m in multiples))


def concat(values, axis):
    return np.concatenate(list(values), axis=axis)


def transpose(x, perm=None):
    return np.transpose(x, perm)


def reshape(x, shape):
    return np.reshape(x, shape)


def pad(x, paddings, mode="CONSTANT"):
    return np.pad(x, [(int(a), int(b)) for a, b in np.asarray(paddings)], mode="constant")


def _axis(axis):
    if axis is None:
        return None
    if isinstance(axis, (list, tuple, np.ndarray)):
        return tuple(int(a) for a in axis)
    return int(axis)


def reduce_sum(x, axis=None, keep_dims=False, keepdims=False):
    return np.sum(x, axis=_axis(axis), keepdims=keep_dims or keepdims)


def reduce_mean(x, axis=None, keep_dims=False, keepdims=False):
    return np.mean(x, axis=_axis(axis), keepdims=keep_dims or keepdims)


def reduce_max(x, axis=None, keep_dims=False, keepdims=False):
    return np.max(x, axis=_axis(axis), keepdims=keep_dims or keepdims)


def reduce_min(x, axis=None, keep_dims=False, keepdims=False):
    return np.min(x, axis=_axis(axis), keepdims=keep_dims or keepdims)


def matmul(a, b):
    return np.matmul(a, b)


def tensordot(a, b, axes):
    return np.tensordot(a, b, axes)


def greater(a, b):
    return np.greater(a, b)


def greater_equal(a, b):
    return np.greater_equal(a, b)


def equal(a, b):
    return np.equal(a, b)


def logical_not(a):
    return np.logical_not(a)


def where(cond, x, y):
    return np.where(cond, x, y)


def argmax(x, axis=None, output_type=np.int64):
    return np.argmax(x, axis=axis).astype(output_type)


def one_hot(indices, depth, dtype=np.float32):
    """[TF-sem] out-of-range -> all zeros; in-range index -> 1.0 at that position."""
    idx = np.asarray(indices).astype(np.int64)
    out = np.zeros(idx.shape + (int(depth),), dtype=dtype)
    ok = (idx >= 0) & (idx < depth)
    np.put_along_axis(out, np.where(ok, idx, 0)[..., None], ok[..., None].astype(dtype), axis=-1)
    return out


# ---- tf.nn -----------------------------------------------------------------------------
nn = types.ModuleType("tensorflow.nn")


def _relu(x):
    return np.maximum(x, 0)


def _l2_normalize(x, dim=None, axis=None, epsilon=1e-12):
    ax = _axis(dim if dim is not None else axis)
    ss = np.sum(np.square(x), axis=ax, keepdims=True)
    return x / np.sqrt(np.maximum(ss, epsilon))


def _softmax_xent(labels=None, logits=None, dim=-1, **_):
    """[TF-sem] -sum(labels * log_softmax(logits)) along the last axis."""
    z = logits
    m = np.max(z, axis=-1, keepdims=True)
    lse = m + np.log(np.sum(np.exp(z - m), axis=-1, keepdims=True))
    return np.sum(labels * (lse - z), axis=-1)


def _top_k(x, k=1, sorted=True):  # noqa: A002
    """[TF-sem] descending values, ties broken towards the lower index."""
    k = int(k)
    n = x.shape[-1]
    flat = np.asarray(x).reshape(-1, n)
    idx = np.empty((flat.shape[0], k), dtype=np.int32)
    ar = np.arange(n)
    for r in range(flat.shape[0]):
        idx[r] = np.lexsort((ar, -flat[r].astype(np.float64)))[:k]
    val = np.take_along_axis(flat, idx.astype(np.int64), 1)
    return _t(val.reshape(x.shape[:-1] + (k,))), _t(idx.reshape(x.shape[:-1] + (k,)))


def _bias_add(x, b, data_format=None):
    return x + b


nn.relu = _relu
nn.l2_normalize = _l2_normalize
nn.softmax_cross_entropy_with_logits = _softmax_xent
nn.top_k = _top_k
nn.bias_add = _bias_add
nn.sigmoid = sigmoid
nn.tanh = np.tanh
nn.dropout = lambda x, keep_prob, noise_shape=None, **_: x
tanh = np.tanh

# ---- tf.layers -------------------------------------------------------------------------
layers = types.ModuleType("tensorflow.layers")


def _snake(name):
    s1 = re.sub("(.)([A-Z][a-z0-9]+)", r"\1_\2", name)
    return re.sub("([a-z])([A-Z0-9])", r"\1_\2", s1).lower()


class Layer(object):
    """[TF-sem] tf.layers.Layer: __call__ opens a variable scope named after the layer
    (given name, else the snake-cased class name), builds on first use, then calls."""

    def __init__(self, trainable=True, name=None, dtype=None, activity_regularizer=None, **kwargs):
        self.trainable = trainable
        self.name = name
        self.dtype = dtype
        self.built = False

    def build(self, input_shape):
        self.built = True

    def add_variable(self, name, shape=None, **kw):
        return get_variable(name, shape=shape)

    def __call__(self, inputs, *args, **kwargs):
        scope = self.name if self.name is not None else _snake(type(self).__name__)
        with variable_scope(scope):
            if not self.built:
                self.build(list(np.shape(inputs)))
                self.built = True
            return self.call(inputs, *args, **kwargs)


class Dropout(Layer):
    """[TF-sem] identity when rate == 0 or not training; the hot path runs at rate 0.0 (args.py:64)."""

    def __init__(self, rate=0.5, noise_shape=None, seed=None, name=None, **kw):
        super(Dropout, self).__init__(name=name)
        self.rate = rate

    def __call__(self, inputs, training=False):
        if float(self.rate) != 0.0 and np.asarray(training).any():
            raise NotImplementedError("dropout > 0 is outside the oracle's scope")
        return inputs


layers.Layer = Layer
layers.Dropout = Dropout

# ---- tf.contrib ------------------------------------------------------------------------
from . import contrib  # noqa: E402,F401

sys.modules[__name__ + ".nn"] = nn
sys.modules[__name__ + ".layers"] = layers
