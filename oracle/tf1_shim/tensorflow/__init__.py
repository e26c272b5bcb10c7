"""Eager NumPy stand-in for the TF-1.0/1.1 calls made by the reference hot path.

TEST INFRASTRUCTURE (see ../README.md).  Restates published TensorFlow 1.x op
semantics; used only by oracle/make_golden.py to execute the reference's own
source files.  The working precision is module-global (`set_precision`) so the
same reference code can be run in float64 (ground truth) or float32.
"""
import contextlib

import numpy as np

_PREC = {"dtype": np.float64}
float32 = "float32"   # symbolic; mapped to the working precision
int32 = np.int32


def set_precision(dt):
    _PREC["dtype"] = np.dtype(dt).type


def _wd():
    return _PREC["dtype"]


# ------------------------------------------------------------------ tensors --
class Dimension(object):
    def __init__(self, v):
        self.value = v

    def __int__(self):
        return int(self.value)

    def __eq__(self, o):
        return self.value == (o.value if isinstance(o, Dimension) else o)


class TensorShape(object):
    def __init__(self, dims):
        self._d = list(dims)
        self.ndims = len(self._d)

    def as_list(self):
        return list(self._d)

    def __getitem__(self, i):
        return Dimension(self._d[i])

    def __len__(self):
        return len(self._d)

    def __iter__(self):
        return iter(Dimension(d) for d in self._d)


class Tensor(object):
    __array_priority__ = 1000

    def __init__(self, a):
        if isinstance(a, Tensor):
            a = a.a
        a = np.asarray(a)
        if a.dtype.kind == "f":
            a = a.astype(_wd(), copy=False)
        self.a = a

    @property
    def dtype(self):
        return float32 if self.a.dtype.kind == "f" else self.a.dtype

    def get_shape(self):
        return TensorShape(self.a.shape)

    @property
    def shape(self):
        return TensorShape(self.a.shape)

    def eval(self):
        return self.a

    def __getitem__(self, idx):
        return Tensor(self.a[idx])

    def _b(op):
        def f(self, o):
            return Tensor(op(self.a, _np(o)))
        return f

    def _r(op):
        def f(self, o):
            return Tensor(op(_np(o), self.a))
        return f

    __add__ = _b(np.add); __radd__ = _r(np.add)
    __sub__ = _b(np.subtract); __rsub__ = _r(np.subtract)
    __mul__ = _b(np.multiply); __rmul__ = _r(np.multiply)
    __truediv__ = _b(np.true_divide); __rtruediv__ = _r(np.true_divide)
    __lt__ = _b(np.less)

    def __neg__(self):
        return Tensor(-self.a)

    def __bool__(self):
        return bool(self.a)


def _np(x):
    if isinstance(x, Tensor):
        return x.a
    a = np.asarray(x)
    if a.dtype.kind == "f":
        a = a.astype(_wd())
    return a


def _t(x):
    return Tensor(x)


# --------------------------------------------------------- variable scopes --
class _Scope(object):
    def __init__(self, name, initializer=None):
        self.name = name
        self.initializer = initializer

    def set_partitioner(self, p):
        pass


class _Store(object):
    def __init__(self):
        self.vars = {}        # values supplied by the caller (a "checkpoint")
        self.created = {}     # every variable the reference asked for
        self.stack = [_Scope("")]
        self.rng = np.random.RandomState(0)


_S = _Store()


def reset_store(values=None, seed=0):
    _S.vars = dict(values or {})
    _S.created = {}
    _S.stack = [_Scope("")]
    _S.rng = np.random.RandomState(seed)


def created_variables():
    return dict(_S.created)


@contextlib.contextmanager
def variable_scope(name_or_scope, initializer=None, reuse=None, **kw):
    cur = _S.stack[-1]
    if isinstance(name_or_scope, _Scope):
        new = _Scope(name_or_scope.name, initializer or name_or_scope.initializer)
    else:
        full = (cur.name + "/" + name_or_scope) if cur.name else name_or_scope
        new = _Scope(full, initializer or cur.initializer)
    _S.stack.append(new)
    try:
        yield new
    finally:
        _S.stack.pop()


def get_variable_scope():
    return _S.stack[-1]


def get_variable(name, shape=None, dtype=None, initializer=None, **kw):
    cur = _S.stack[-1]
    full = (cur.name + "/" + name) if cur.name else name
    shape = tuple(int(d) for d in shape)
    if full in _S.vars:
        v = np.asarray(_S.vars[full])
        if tuple(v.shape) != shape:
            raise ValueError("variable %s: stored shape %s, requested %s" % (full, v.shape, shape))
    else:
        init = initializer or cur.initializer
        if init is None:
            raise ValueError("variable %s has no stored value and no initializer" % full)
        v = init(shape, _S.rng)
    _S.created[full] = np.asarray(v, np.float32)
    return Tensor(np.asarray(v, np.float32))


def random_uniform_initializer(minval=0.0, maxval=1.0, seed=None, dtype=None):
    return lambda shape, rng: rng.uniform(minval, maxval, size=shape).astype(np.float32)


def constant_initializer(value=0.0, dtype=None):
    return lambda shape, rng: np.full(shape, value, np.float32)


def placeholder(dtype=None, shape=None, name=None):
    return Tensor(np.zeros([int(d) for d in shape]))


# --------------------------------------------------------------------- ops --
def constant(v, dtype=None, shape=None, name=None):
    a = np.asarray(v, dtype if dtype not in (None, float32) else None)
    if shape is not None:
        a = np.broadcast_to(a, shape).copy()
    return Tensor(a)


def reshape(x, shape, name=None):
    return Tensor(np.reshape(_np(x), [int(d) for d in shape]))


def concat(values, axis, name=None):
    return Tensor(np.concatenate([_np(v) for v in values], axis))


def stack(values, axis=0, name=None):
    return Tensor(np.stack([_np(v) for v in values], axis))


def transpose(x, perm=None, name=None):
    return Tensor(np.transpose(_np(x), perm))


def expand_dims(x, axis, name=None):
    return Tensor(np.expand_dims(_np(x), axis))


def squeeze(x, axis=None, name=None):
    return Tensor(np.squeeze(_np(x), axis))


def slice(x, begin, size, name=None):  # noqa: A001 (TF name)
    a = _np(x)
    idx = []
    for b, s, d in zip(begin, size, a.shape):
        idx.append(np.s_[b:(d if s == -1 else b + s)])
    return Tensor(a[tuple(idx)])


def tanh(x, name=None):
    return Tensor(np.tanh(_np(x)))


def sigmoid(x, name=None):
    return Tensor(1.0 / (1.0 + np.exp(-_np(x))))


def multiply(a, b, name=None):
    return Tensor(_np(a) * _np(b))


def add(a, b, name=None):
    return Tensor(_np(a) + _np(b))


def add_n(xs, name=None):
    acc = _np(xs[0])
    for v in xs[1:]:
        acc = acc + _np(v)
    return Tensor(acc)


def div(a, b, name=None):
    return Tensor(_np(a) / _np(b))


def pow(a, b, name=None):  # noqa: A001
    return Tensor(np.power(_np(a), _np(b)))


def reduce_sum(x, axis=None, keep_dims=False, name=None):
    return Tensor(np.sum(_np(x), axis=axis, keepdims=keep_dims))


def reduce_prod(x, axis=None, keep_dims=False, name=None):
    return Tensor(np.prod(_np(x), axis=axis, keepdims=keep_dims))


def matmul(a, b, name=None):
    return Tensor(np.matmul(_np(a), _np(b)))


def zeros(shape, dtype=None, name=None):
    return Tensor(np.zeros([int(d) for d in shape]))


def zeros_like(x, name=None):
    return Tensor(np.zeros_like(_np(x)))


def ones(shape, dtype=None, name=None):
    return Tensor(np.ones([int(d) for d in shape]))


class _NN(object):
    @staticmethod
    def softplus(x, name=None):
        return Tensor(np.logaddexp(0.0, _np(x)))

    @staticmethod
    def softmax(x, dim=-1, name=None):
        a = _np(x)
        e = np.exp(a - np.max(a, axis=dim, keepdims=True))
        return Tensor(e / np.sum(e, axis=dim, keepdims=True))

    @staticmethod
    def l2_normalize(x, dim, epsilon=1e-12, name=None):
        a = _np(x)
        ss = np.sum(np.square(a), axis=dim, keepdims=True)
        return Tensor(a * (1.0 / np.sqrt(np.maximum(ss, epsilon))))

    @staticmethod
    def bias_add(x, b, name=None):
        return Tensor(_np(x) + _np(b))


nn = _NN()


# ---------------------------------------------------- TensorArray / while --
class TensorArray(object):
    def __init__(self, dtype=None, size=None, name=None, **kw):
        self._v = [None] * int(size)

    def write(self, index, value):
        self._v[int(_np(index))] = Tensor(value)
        return self

    def read(self, index):
        return self._v[int(_np(index))]

    def unstack(self, value, name=None):
        a = _np(value)
        self._v = [Tensor(a[i]) for i in range(a.shape[0])]
        return self

    def stack(self, name=None):
        return Tensor(np.stack([_np(v) for v in self._v], 0))


def while_loop(cond, body, loop_vars, **kw):
    vs_ = tuple(loop_vars)
    while bool(_np(cond(*vs_))):
        vs_ = tuple(body(*vs_))
    return vs_


class _TestCase(object):
    pass


class _Test(object):
    TestCase = _TestCase

    @staticmethod
    def main():
        pass


test = _Test()

from tensorflow import contrib  # noqa: E402,F401
from tensorflow import python   # noqa: E402,F401
