"""NumPy-backed stand-in for `mlx.core` (TEST INFRASTRUCTURE ONLY, see package docstring).

Semantics assumed for MLX 0.7.0 (third-party, source not under /root/reference):
  * default float dtype is float32; float64 inputs are narrowed to float32;
  * `linspace(a, b, n)` = `arange(n, f32) * f32((b - a) / (n - 1)) + f32(a)`;
  * every elementwise op rounds to fp32 once (NumPy fp32 ufuncs do the same).
"""
import numpy as _np

float32 = _np.float32
int32 = _np.int32
int64 = _np.int64
uint32 = _np.uint32
pi = _np.pi
cpu = "cpu"
gpu = "gpu"


def _c(x):
    keep = isinstance(x, mxarray)  # MLX-promotion arrays (see `mxarray`) stay what they are
    x = _np.asarray(x)
    if x.dtype == _np.float64:
        x = x.astype(_np.float32)
    return x.view(mxarray) if keep else x


def array(x, dtype=None):
    if dtype is not None:
        return _np.asarray(x).astype(dtype)
    a = _np.asarray(x)
    if a.dtype == _np.float64:
        a = a.astype(_np.float32)
    elif a.dtype == _np.int64 and not isinstance(x, _np.ndarray):
        a = a.astype(_np.int32)
    return a


def set_default_device(_):
    return None


def eval(*_a, **_k):
    return None


def compile(fn=None, inputs=None, outputs=None):
    if fn is None:
        return lambda f: f
    return fn


def linspace(start, stop, num=50, dtype=float32):
    if num == 1:  # mx.linspace(a, b, num=1) is [a]
        return _np.array([start], dtype=_np.float32).astype(dtype)
    seq = _np.arange(num, dtype=_np.float32)
    step = _np.float32((float(stop) - float(start)) / (num - 1))
    return (seq * step + _np.float32(start)).astype(dtype)


class mxarray(_np.ndarray):
    """ndarray with MLX's type promotion in ufuncs: an integer array meeting a floating array (or scalar) is
    promoted to float32 (NumPy would go to float64), and float64 never appears.  Returned by `arange` only (its one
    caller on the path is MultiHashEncoding.__init__, multi_hash.py:32-40: `growing_factor ** levels`)."""

    def __array_ufunc__(self, ufunc, method, *inputs, out=None, **kwargs):
        arrs = [(_np.asarray(x) if isinstance(x, _np.ndarray) else x) for x in inputs]
        has_float = any((isinstance(a, float) or (isinstance(a, _np.ndarray) and a.dtype.kind == "f")) for a in arrs)
        if has_float:
            arrs = [a.astype(_np.float32) if isinstance(a, _np.ndarray) and a.dtype != _np.float32 else a for a in arrs]
        r = getattr(ufunc, method)(*arrs, **kwargs)
        if isinstance(r, _np.ndarray):
            if r.dtype == _np.float64:
                r = r.astype(_np.float32)
            return r.view(mxarray)
        return r


def arange(*a, dtype=None):
    r = _np.arange(*a)
    if dtype is not None:
        return r.astype(dtype).view(mxarray)
    if r.dtype == _np.int64:
        return r.astype(_np.int32).view(mxarray)
    return _c(r).view(mxarray)


def concatenate(arrs, axis=0):
    return _c(_np.concatenate([_np.asarray(a) for a in arrs], axis=axis))


def stack(arrs, axis=0):
    return _c(_np.stack([_np.asarray(a) for a in arrs], axis=axis))


def reshape(a, shape):
    return _np.reshape(a, tuple(shape))


def repeat(a, repeats, axis=None):
    return _np.repeat(a, repeats, axis=axis)


def expand_dims(a, axis):
    return _np.expand_dims(a, axis)


def split(a, indices_or_sections, axis=0):
    return _np.split(a, indices_or_sections, axis=axis)


def flatten(a, start_axis=0, end_axis=-1):
    a = _np.asarray(a)
    nd = a.ndim
    s = start_axis % nd
    e = end_axis % nd
    shape = a.shape[:s] + (-1,) + a.shape[e + 1:]
    return a.reshape(shape)


def take(a, indices, axis=None):
    return _np.take(a, indices, axis=axis)


def sort(a, axis=-1):
    return _np.sort(a, axis=axis, kind="stable")


def sum(a, axis=None, keepdims=False):
    return _c(_np.sum(a, axis=axis, keepdims=keepdims, dtype=_np.asarray(a).dtype))


def mean(a, axis=None, keepdims=False):
    return _c(_np.mean(a, axis=axis, keepdims=keepdims, dtype=_np.asarray(a).dtype))


def cumsum(a, axis=None):
    a = _np.asarray(a)
    return _np.cumsum(a, axis=axis, dtype=a.dtype)


def exp(a):
    return _c(_np.exp(_c(a)))


def log(a):
    return _c(_np.log(_np.asarray(a, dtype=_np.float32)))


def log10(a):
    return _c(_np.log10(_np.asarray(a, dtype=_np.float32)))


def sin(a):
    return _c(_np.sin(_c(a)))


def cos(a):
    return _c(_np.cos(_c(a)))


def floor(a):
    return _c(_np.floor(a))


def ceil(a):
    return _c(_np.ceil(a))


def square(a):
    return _c(_np.square(a))


def sqrt(a):
    return _c(_np.sqrt(a))


def clip(a, a_min=None, a_max=None):
    return _c(_np.clip(a, a_min, a_max))


def minimum(a, b):
    return _c(_np.minimum(a, b))


def maximum(a, b):
    return _c(_np.maximum(a, b))


def where(c, a, b):
    return _c(_np.where(c, a, b))


def ones_like(a):
    return _np.ones_like(a)


def zeros_like(a):
    return _np.zeros_like(a)


def zeros(shape, dtype=float32):
    return _np.zeros(tuple(shape) if not isinstance(shape, int) else shape, dtype=dtype)


def ones(shape, dtype=float32):
    return _np.ones(tuple(shape) if not isinstance(shape, int) else shape, dtype=dtype)


class _Linalg:
    @staticmethod
    def norm(a, axis=None, keepdims=False):
        a = _np.asarray(a)
        return _c(_np.sqrt(_np.sum(a * a, axis=axis, keepdims=keepdims, dtype=a.dtype)))


linalg = _Linalg()


class _Random:
    """Seedable RNG so golden generation is reproducible; the hot path never relies on MLX's RNG
    stream (random tensors are explicit inputs in the oracle and the kernels)."""

    def __init__(self):
        self._g = _np.random.default_rng(0)

    def seed(self, s):
        self._g = _np.random.default_rng(s)

    def uniform(self, low=0.0, high=1.0, shape=()):
        return (self._g.random(size=tuple(shape), dtype=_np.float32) * _np.float32(high - low)
                + _np.float32(low)).astype(_np.float32)

    def normal(self, shape=()):
        return self._g.standard_normal(size=tuple(shape), dtype=_np.float32)


random = _Random()
