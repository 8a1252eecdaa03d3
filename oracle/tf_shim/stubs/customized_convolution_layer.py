"""numpy stand-in for the reference's customized_convolution_layer.py (a copy of TF
layers/convolutional.py).  TEST INFRASTRUCTURE.  Restates _Conv.build/call for rank 1
(customized_convolution_layer.py:127-198) and Conv1D (:230-319): kernel [K,Cin,Cout],
cross-correlation, VALID padding, dilation, bias_add, activation."""
import numpy as np

import tensorflow as tf


class Conv1D(tf.layers.Layer):
    def __init__(self, filters, kernel_size, strides=1, padding="valid", data_format="channels_last",
                 dilation_rate=1, activation=None, use_bias=True, kernel_initializer=None,
                 bias_initializer=None, kernel_regularizer=None, bias_regularizer=None,
                 activity_regularizer=None, kernel_constraint=None, bias_constraint=None,
                 trainable=True, has_weightnorm=False, name=None, **kwargs):
        super(Conv1D, self).__init__(trainable=trainable, name=name, **kwargs)
        as_tuple = lambda v: tuple(v) if isinstance(v, (list, tuple)) else (int(v),)  # noqa: E731
        self.rank = 1
        self.filters = int(filters)
        self.kernel_size = as_tuple(kernel_size)
        self.strides = as_tuple(strides)
        self.padding = padding
        self.data_format = data_format
        self.dilation_rate = as_tuple(dilation_rate)
        self.activation = activation
        self.use_bias = use_bias
        self.has_weightnorm = has_weightnorm
        assert padding == "valid" and data_format == "channels_last" and self.strides == (1,)

    def build(self, input_shape):
        input_dim = input_shape[-1]
        self.kernel = self.add_variable("kernel", shape=self.kernel_size + (input_dim, self.filters))
        if self.has_weightnorm:
            g = tf.get_variable("g", shape=[self.filters])
            self.kernel = np.reshape(g, [1, 1, self.filters]) * tf.nn.l2_normalize(self.kernel, [0, 1])
        self.bias = self.add_variable("bias", shape=(self.filters,)) if self.use_bias else None
        self.built = True

    def call(self, inputs):
        x = np.asarray(inputs)
        K, d = self.kernel_size[0], self.dilation_rate[0]
        L_out = x.shape[1] - (K - 1) * d
        out = np.zeros((x.shape[0], L_out, self.filters), dtype=np.result_type(x, self.kernel))
        for k in range(K):                       # [TF-sem] nn_ops.Convolution == cross-correlation
            out += np.matmul(x[:, k * d:k * d + L_out, :], self.kernel[k])
        if self.use_bias:
            out = out + self.bias
        if self.activation is not None:
            return self.activation(out)
        return out


Conv2D = Conv3D = None
