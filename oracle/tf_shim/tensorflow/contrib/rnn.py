"""tf.contrib.rnn.GRUCell / MultiRNNCell on numpy (TEST INFRASTRUCTURE).

The stock TF-1.6 cells are not in the reference tree, but the reference vendors their source:
customed_gru_cell.py:273-337 (GRUCell), :998-1073 (MultiRNNCell), :1121-1197 (_Linear).  This
file restates those three, variable names included ('gates'/'candidate' + 'kernel'/'bias').
"""
import numpy as np

import tensorflow as tf


def _linear(args, scope, out):
    """customed_gru_cell.py:1187-1197: matmul(concat(args, 1), kernel) + bias."""
    with tf.variable_scope(scope):
        x = np.concatenate(args, 1)
        w = tf.get_variable("kernel", [x.shape[1], out])
        b = tf.get_variable("bias", [out])
    return np.matmul(x, w) + b


class GRUCell(object):
    def __init__(self, num_units, activation=None, reuse=None, kernel_initializer=None, bias_initializer=None):
        self._num_units = num_units
        self._activation = activation or np.tanh

    @property
    def state_size(self):
        return self._num_units

    def __call__(self, inputs, state):
        # RNNCell.__call__ -> Layer.__call__ opens scope 'gru_cell' [TF-sem]; body = GRUCell.call (:309-337)
        with tf.variable_scope("gru_cell"):
            value = tf.sigmoid(_linear([inputs, state], "gates", 2 * self._num_units))
            r, u = np.split(value, 2, axis=1)
            r_state = r * state
            c = self._activation(_linear([inputs, r_state], "candidate", self._num_units))
            new_h = u * state + (1 - u) * c
        return new_h, new_h


class MultiRNNCell(object):
    def __init__(self, cells, state_is_tuple=True):
        self._cells = cells
        self._state_is_tuple = state_is_tuple

    def __call__(self, inputs, state):
        # customed_gru_cell.py:1050-1073 under scope 'multi_rnn_cell'
        assert not self._state_is_tuple
        with tf.variable_scope("multi_rnn_cell"):
            pos = 0
            cur = inputs
            new_states = []
            for i, cell in enumerate(self._cells):
                with tf.variable_scope("cell_%d" % i):
                    cur_state = state[:, pos:pos + cell.state_size]
                    pos += cell.state_size
                    cur, ns = cell(cur, cur_state)
                    new_states.append(ns)
        return cur, np.concatenate(new_states, 1)
