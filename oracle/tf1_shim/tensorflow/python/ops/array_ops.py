"""array_ops.split / concat stand-ins (TF 1.x: split(value, size_splits, axis))."""
import numpy as np

import tensorflow as tf

concat = tf.concat


def split(value, num_or_size_splits, axis=0, name=None):
    a = tf._np(value)
    if isinstance(num_or_size_splits, int):
        return [tf.Tensor(p) for p in np.split(a, num_or_size_splits, axis)]
    offs = np.cumsum(num_or_size_splits)[:-1]
    return [tf.Tensor(p) for p in np.split(a, offs, axis)]
