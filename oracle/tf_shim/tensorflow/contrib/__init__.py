"""tf.contrib stand-in: only ``rnn`` (GRUCell, MultiRNNCell) is on the hot path."""
import types

from . import rnn  # noqa: F401

layers = types.SimpleNamespace(
    layer_norm=lambda *a, **k: (_ for _ in ()).throw(NotImplementedError("has_layernorm is off")),
    fully_connected=lambda *a, **k: (_ for _ in ()).throw(NotImplementedError()),
)
