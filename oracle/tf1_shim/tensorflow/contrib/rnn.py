"""tf.contrib.rnn.BasicLSTMCell / MultiRNNCell, TF 1.0/1.1 semantics, restated.

Published algorithm (tensorflow/contrib/rnn/python/ops/core_rnn_cell_impl.py,
TF 1.0): non-tuple state = concat([c, h], 1);
concat = _linear([inputs, h], 4*num_units, bias=True) with variables
"weights"/"biases" created in scope "basic_lstm_cell";
i, j, f, o = split(concat, 4, axis=1);
new_c = c*sigmoid(f + forget_bias) + sigmoid(i)*tanh(j); new_h = tanh(new_c)*sigmoid(o).
MultiRNNCell: scope (given or "multi_rnn_cell") / "cell_%d"; per-layer state is
the slice [pos, pos + state_size) of the flat state; outputs concat of new states.
TEST INFRASTRUCTURE.
"""
import numpy as np

import tensorflow as tf


class BasicLSTMCell(object):
    def __init__(self, num_units, forget_bias=1.0, state_is_tuple=True, **kw):
        assert not state_is_tuple, "shim implements the non-tuple path the reference uses"
        self._n = num_units
        self._fb = forget_bias

    @property
    def state_size(self):
        return 2 * self._n

    @property
    def output_size(self):
        return self._n

    def zero_state(self, batch_size, dtype):
        return tf.Tensor(np.zeros((batch_size, self.state_size)))

    def __call__(self, inputs, state, scope=None):
        with tf.variable_scope(scope or "basic_lstm_cell"):
            x, s = tf._np(inputs), tf._np(state)
            c, h = s[:, :self._n], s[:, self._n:]
            args = np.concatenate([x, h], 1)
            w = tf._np(tf.get_variable("weights", [args.shape[1], 4 * self._n]))
            b = tf._np(tf.get_variable("biases", [4 * self._n],
                                       initializer=tf.constant_initializer(0.0)))
            z = args @ w + b
            i, j, f, o = np.split(z, 4, axis=1)
            sig = lambda v: 1.0 / (1.0 + np.exp(-v))
            new_c = c * sig(f + self._fb) + sig(i) * np.tanh(j)
            new_h = np.tanh(new_c) * sig(o)
            return tf.Tensor(new_h), tf.Tensor(np.concatenate([new_c, new_h], 1))


class MultiRNNCell(object):
    def __init__(self, cells, state_is_tuple=True):
        assert not state_is_tuple
        self._cells = list(cells)

    @property
    def state_size(self):
        return sum(c.state_size for c in self._cells)

    def zero_state(self, batch_size, dtype):
        return tf.Tensor(np.zeros((batch_size, self.state_size)))

    def __call__(self, inputs, state, scope=None):
        with tf.variable_scope(scope or "multi_rnn_cell"):
            pos = 0
            cur = inputs
            s = tf._np(state)
            new_states = []
            for i, cell in enumerate(self._cells):
                with tf.variable_scope("cell_%d" % i):
                    cs = tf.Tensor(s[:, pos:pos + cell.state_size])
                    pos += cell.state_size
                    cur, ns = cell(cur, cs)
                    new_states.append(tf._np(ns))
        return cur, tf.Tensor(np.concatenate(new_states, 1))
