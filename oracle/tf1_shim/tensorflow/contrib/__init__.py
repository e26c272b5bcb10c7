"""tf.contrib stand-in: only tf.contrib.rnn (see ../../README.md)."""
from tensorflow.contrib import rnn  # noqa: F401
