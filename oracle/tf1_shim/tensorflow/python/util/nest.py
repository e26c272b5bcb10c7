"""tensorflow.python.util.nest.is_sequence stand-in."""


def is_sequence(x):
    return isinstance(x, (list, tuple))
