"""ORACLE SUPPORT (test infrastructure, NOT product code, NOT TensorFlow).

An eager float32 numpy stand-in for exactly the TensorFlow ~0.12 Python API
surface that the reference's `nms_net/network.py` (class Gnet) touches, so that
the reference's OWN model code can be imported from /root/reference and
executed here, unmodified, to produce golden vectors
(oracle/run_reference_graph.py -> tests/golden/).  Semantics follow the TF 0.12
documentation of each op (argument order included: `tf.concat(dim, values)`,
`tf.select`, `tf.pack`, `tf.mul`, `tf.sub`, ...).  Every op evaluates
immediately on numpy arrays; float math stays float32 (numpy keeps float32 for
float32-array (op) python-scalar, like TF's constant conversion).

Variables: `tf.contrib.layers.fully_connected` looks its `weights`/`biases` up
by variable-scope name in `PARAMS` (a dict the driver fills from the shared
seeded generator); a missing name raises.
"""
import contextlib

import numpy as np

float32 = np.float32
float64 = np.float64
int32 = np.int32
int64 = np.int64
bool = np.bool_  # noqa: A001  (the reference writes tf.bool)

PARAMS = {}            # scope/name -> np.ndarray, set by the driver
OP_LIBRARIES = {}      # basename of the .so -> python object standing in for it
_scope = []
_created = []
_losses = []


def reset():
    del _scope[:]
    del _created[:]
    del _losses[:]
    _default_names.clear()


_default_names = {}     # (enclosing scope, default name) -> uses so far


def unique_default_scope(default_name):
    """tf.variable_scope(None, default_name=...): `fully_connected`, then
    `fully_connected_1`, ... within one enclosing scope."""
    key = (current_scope(), default_name)
    n = _default_names.get(key, 0)
    _default_names[key] = n + 1
    return default_name if n == 0 else '%s_%d' % (default_name, n)


class T(object):
    """Eager tensor: a numpy array with TF-flavoured operators."""
    __array_priority__ = 1000

    def __init__(self, value, dtype=None):
        if isinstance(value, T):
            value = value.v
        self.v = np.asarray(value, dtype=dtype)

    # -- TF tensor API used by the reference --
    @property
    def dtype(self):
        return self.v.dtype.type

    def set_shape(self, shape):
        shape = list(shape)
        assert len(shape) == self.v.ndim, (shape, self.v.shape)
        for want, have in zip(shape, self.v.shape):
            assert want is None or want == have, (shape, self.v.shape)

    def get_shape(self):
        return self.v.shape

    @property
    def name(self):
        return 'eager:0'

    def __getitem__(self, idx):
        if isinstance(idx, T):
            idx = idx.v
        return T(self.v[idx])

    def __iter__(self):
        raise TypeError('tensors are not iterable')

    def __bool__(self):
        return builtins_bool(self.v)

    __nonzero__ = __bool__

    def _bin(self, other, fn, rev=False):
        o = other.v if isinstance(other, T) else other
        if not isinstance(o, np.ndarray) and np.issubdtype(self.v.dtype, np.floating):
            o = self.v.dtype.type(o)   # TF converts python scalars to the tensor dtype
        return T(fn(o, self.v) if rev else fn(self.v, o))

    def __add__(self, o): return self._bin(o, np.add)
    def __radd__(self, o): return self._bin(o, np.add, True)
    def __sub__(self, o): return self._bin(o, np.subtract)
    def __rsub__(self, o): return self._bin(o, np.subtract, True)
    def __mul__(self, o): return self._bin(o, np.multiply)
    def __rmul__(self, o): return self._bin(o, np.multiply, True)
    def __truediv__(self, o): return self._bin(o, np.divide)
    def __rtruediv__(self, o): return self._bin(o, np.divide, True)
    __div__ = __truediv__
    def __pow__(self, o): return self._bin(o, np.power)
    def __neg__(self): return T(-self.v)
    def __ge__(self, o): return self._bin(o, np.greater_equal)
    def __gt__(self, o): return self._bin(o, np.greater)
    def __le__(self, o): return self._bin(o, np.less_equal)
    def __lt__(self, o): return self._bin(o, np.less)


import builtins  # noqa: E402
builtins_bool = builtins.bool


def _v(x):
    return x.v if isinstance(x, T) else x


def _f(x, like=None):
    """Operand -> numpy, python scalars taking the dtype of `like`."""
    x = _v(x)
    if like is not None and not isinstance(x, np.ndarray):
        return _v(like).dtype.type(x)
    return np.asarray(x)


# ------------------------------------------------------------------ graph-ish
def placeholder(dtype, shape=None, name=None):
    raise RuntimeError('the eager stand-in has no placeholders: pass batch= to Gnet')


@contextlib.contextmanager
def variable_scope(name, reuse=None, **kwargs):
    _scope.append(name)
    try:
        yield name
    finally:
        _scope.pop()


name_scope = variable_scope


def current_scope():
    return '/'.join(_scope)


class Variable(object):
    def __init__(self, name, value):
        self.name = name + ':0'
        self.value = value

    class _Op(object):
        def __init__(self, name):
            self.name = name

    @property
    def op(self):
        return Variable._Op(self.name[:-2])


def get_param(name):
    if name not in PARAMS:
        raise KeyError('variable %r not provided to the TF stand-in' % name)
    if name not in [v.name[:-2] for v in _created]:
        _created.append(Variable(name, PARAMS[name]))
    return PARAMS[name]


def trainable_variables():
    return list(_created)


def constant_initializer(value=0.0, dtype=float32):
    return ('constant', value)


def constant(value, dtype=None, shape=None, name=None):
    return T(np.asarray(value, dtype=dtype))


def convert_to_tensor(value, dtype=None):
    return T(value, dtype=dtype)


def stop_gradient(x, name=None):
    return x


def NotDifferentiable(name):
    pass


def RegisterShape(name):
    return lambda fn: fn


class _OpLib(object):
    pass


def load_op_library(path):
    import os.path
    base = os.path.basename(path)
    if base not in OP_LIBRARIES:
        raise ImportError('no stand-in registered for op library %s' % base)
    return OP_LIBRARIES[base]


# ------------------------------------------------------------------- array ops
def shape(x, name=None):
    return T(np.asarray(np.shape(_v(x)), dtype=np.int32))


def reshape(x, shp, name=None):
    shp = _v(shp)
    return T(np.reshape(_v(x), [int(s) for s in np.asarray(shp).reshape(-1)]))


def slice(x, begin, size, name=None):  # noqa: A001
    x = _v(x)
    idx = tuple(builtins.slice(b, None if s == -1 else b + s) for b, s in zip(begin, size))
    return T(x[idx])


def pack(values, axis=0, name=None):
    return T(np.stack([np.asarray(_v(v)) for v in values], axis=axis))


stack = pack


def concat(concat_dim, values, name=None):
    """TF <= 0.12 argument order: (concat_dim, values)."""
    assert isinstance(concat_dim, int), 'tf.concat(dim, values) order expected'
    return T(np.concatenate([_v(v) for v in values], axis=concat_dim))


def tile(x, multiples, name=None):
    return T(np.tile(_v(x), [int(m) for m in np.asarray(_v(multiples)).reshape(-1)]))


def expand_dims(x, dim, name=None):
    return T(np.expand_dims(_v(x), dim))


def zeros(shp, dtype=float32, name=None):
    return T(np.zeros([int(s) for s in np.asarray(_v(shp)).reshape(-1)], dtype=dtype))


def zeros_like(x, dtype=None):
    return T(np.zeros_like(_v(x), dtype=dtype))


def range(*args, **kwargs):  # noqa: A001
    return T(np.arange(*[int(_v(a)) for a in args], dtype=np.int32))


def cast(x, dtype, name=None):
    return T(_v(x).astype(dtype))


def gather(params, indices, name=None):
    return T(_v(params)[_v(indices)])


def gather_nd(params, indices, name=None):
    idx = _v(indices)
    return T(_v(params)[tuple(idx[:, i] for i in builtins.range(idx.shape[1]))])


def scatter_nd(indices, updates, shp, name=None):
    out = np.zeros([int(s) for s in np.asarray(_v(shp)).reshape(-1)], dtype=_v(updates).dtype)
    idx = _v(indices)
    np.add.at(out, tuple(idx[:, i] for i in builtins.range(idx.shape[1])), _v(updates))
    return T(out)


def where(cond, name=None):
    """Single-argument form: coordinates of true elements, row-major, int64."""
    return T(np.argwhere(_v(cond)).astype(np.int64))


def select(cond, t, e, name=None):
    """tf.select: a rank-1 condition picks whole rows of higher-rank t / e."""
    c, t, e = _v(cond), _v(t), _v(e)
    if c.ndim == 1 and t.ndim > 1:
        c = c.reshape((-1,) + (1,) * (t.ndim - 1))
    return T(np.where(c, t, e))


def segment_max(data, segment_ids, name=None):
    data, ids = _v(data), _v(segment_ids)
    assert np.all(np.diff(ids) >= 0), 'segment ids must be sorted'
    n = int(ids[-1]) + 1 if ids.size else 0
    out = np.zeros((n,) + data.shape[1:], dtype=data.dtype)  # empty segment -> 0
    if ids.size:
        starts = np.flatnonzero(np.diff(np.concatenate([[-1], ids])) != 0)
        out[ids[starts]] = np.maximum.reduceat(data, starts, axis=0)
    return T(out)


def cond(pred, fn1, fn2, name=None):
    return fn1() if builtins_bool(_v(pred)) else fn2()


# -------------------------------------------------------------------- math ops
def _binary(fn):
    def op(a, b, name=None):
        like = a if isinstance(a, T) else b
        return T(fn(_f(a, like), _f(b, like)))
    return op


add = _binary(np.add)
sub = _binary(np.subtract)
mul = _binary(np.multiply)
div = _binary(np.divide)
maximum = _binary(np.maximum)
minimum = _binary(np.minimum)
equal = _binary(np.equal)
greater_equal = _binary(np.greater_equal)
logical_and = _binary(np.logical_and)


def logical_not(x, name=None):
    return T(np.logical_not(_v(x)))


def sqrt(x, name=None):
    return T(np.sqrt(_v(x)))


def log(x, name=None):
    return T(np.log(_v(x)))


def reduce_sum(x, name=None):
    return T(np.sum(_v(x), dtype=_v(x).dtype))


def reduce_mean(x, name=None):
    return T(np.mean(_v(x), dtype=_v(x).dtype))


def reduce_max(x, name=None):
    return T(np.max(_v(x)))


class _NN(object):
    @staticmethod
    def relu(x, name=None):
        return T(np.maximum(_v(x), _v(x).dtype.type(0)))

    @staticmethod
    def sigmoid_cross_entropy_with_logits(logits, targets, name=None):
        """TF 0.12 positional order (logits, targets):
        max(x,0) - x*z + log(1 + exp(-|x|))."""
        x, z = _v(logits), _v(targets)
        zero = x.dtype.type(0)
        return T(np.maximum(x, zero) - x * z + np.log1p(np.exp(-np.abs(x))))


nn = _NN()

from tensorflow import contrib  # noqa: E402,F401
