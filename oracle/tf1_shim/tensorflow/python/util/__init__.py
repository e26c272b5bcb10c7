from tensorflow.python.util import nest  # noqa: F401
