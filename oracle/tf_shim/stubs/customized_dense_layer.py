"""numpy stand-in for the reference's customized_dense_layer.py (itself a copy of TF layers/core.py).
TEST INFRASTRUCTURE.  Restates Dense.build/call (customized_dense_layer.py:126-172) and the
functional ``dense`` (:184-257).  weight-norm (:140-142) is restated too (pure re-parameterisation)."""
import numpy as np

import tensorflow as tf


class Dense(tf.layers.Layer):
    def __init__(self, units, activation=None, use_bias=True, kernel_initializer=None,
                 bias_initializer=None, kernel_regularizer=None, bias_regularizer=None,
                 activity_regularizer=None, kernel_constraint=None, bias_constraint=None,
                 trainable=True, name=None, has_weightnorm=False, **kwargs):
        super(Dense, self).__init__(trainable=trainable, name=name, **kwargs)
        self.units = int(units)
        self.activation = activation
        self.use_bias = use_bias
        self.has_weightnorm = has_weightnorm

    def build(self, input_shape):
        self.kernel = self.add_variable("kernel", shape=[input_shape[-1], self.units])
        if self.has_weightnorm:
            g = tf.get_variable("g", shape=[self.units])
            self.kernel = g * self.kernel / np.sqrt(np.sum(np.square(self.kernel), 0, keepdims=True))
        self.bias = self.add_variable("bias", shape=[self.units]) if self.use_bias else None
        self.built = True

    def call(self, inputs):
        x = np.asarray(inputs)
        if x.ndim > 2:
            out = np.tensordot(x, self.kernel, [[x.ndim - 1], [0]])
        else:
            out = np.matmul(x, self.kernel)
        if self.use_bias:
            out = out + self.bias
        if self.activation is not None:
            return self.activation(out)
        return out


def dense(inputs, units, activation=None, use_bias=True, kernel_initializer=None, bias_initializer=None,
          kernel_regularizer=None, bias_regularizer=None, activity_regularizer=None,
          kernel_constraint=None, bias_constraint=None, trainable=True, name=None,
          has_weightnorm=False, reuse=None):
    layer = Dense(units, activation=activation, use_bias=use_bias, name=name, has_weightnorm=has_weightnorm)
    return layer(inputs)


class Dropout(tf.layers.Dropout):
    pass
