from tensorflow.python import ops, util  # noqa: F401
