from tensorflow.python.ops import array_ops, init_ops, math_ops, nn_ops, variable_scope  # noqa: F401
