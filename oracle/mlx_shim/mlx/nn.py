"""NumPy-backed stand-in for the slice of `mlx.nn` used by the reference hot path
(TEST INFRASTRUCTURE ONLY).  `Linear` follows MLX 0.7.0: weight [out, in], y = x @ W.T + b,
init U(-1/sqrt(in), +1/sqrt(in)) for weight and bias."""
import math

import numpy as _np


class Module:
    def __init__(self):
        pass

    def parameters(self):
        out = {}
        for k, v in self.__dict__.items():
            if isinstance(v, Module):
                out[k] = v.parameters()
            elif isinstance(v, list) and v and all(isinstance(m, Module) for m in v):
                out[k] = [m.parameters() for m in v]
            elif isinstance(v, _np.ndarray):
                out[k] = v
        return out

    @property
    def state(self):
        return self.parameters()


_rng = _np.random.default_rng(1234)


def seed(s):
    global _rng
    _rng = _np.random.default_rng(s)


class Linear(Module):
    def __init__(self, input_dims, output_dims, bias=True):
        super().__init__()
        scale = math.sqrt(1.0 / input_dims)
        self.weight = _rng.uniform(-scale, scale, size=(output_dims, input_dims)).astype(_np.float32)
        if bias:
            self.bias = _rng.uniform(-scale, scale, size=(output_dims,)).astype(_np.float32)

    def __call__(self, x):
        y = _np.matmul(_np.asarray(x, dtype=_np.float32), self.weight.T)
        if "bias" in self.__dict__:
            y = y + self.bias
        return y.astype(_np.float32)


class Embedding(Module):
    def __init__(self, num_embeddings, dims):
        super().__init__()
        scale = math.sqrt(1.0 / dims)
        self.weight = (_rng.standard_normal(size=(num_embeddings, dims)) * scale).astype(_np.float32)

    def __call__(self, idx):
        return self.weight[idx]


class Identity(Module):
    def __call__(self, x):
        return x


def relu(x):
    return _np.maximum(x, _np.zeros((), dtype=_np.asarray(x).dtype))


class _Init:
    @staticmethod
    def uniform(low=0.0, high=1.0, dtype=_np.float32):
        def f(a):
            return _rng.uniform(low, high, size=a.shape).astype(dtype)
        return f


init = _Init()


def value_and_grad(model, fn):
    raise NotImplementedError("autodiff is not part of the shim; gradients are pinned by torch autograd in oracle/")


class _Init:
    """nn.init.uniform(low, high) -> initialiser returning a NEW array (MLX initialisers are functional; the reference
    discards the result at multi_hash.py:51)."""

    @staticmethod
    def uniform(low=0.0, high=1.0):
        def _f(a):
            return _rng.uniform(low, high, size=_np.shape(a)).astype(_np.float32)
        return _f


init = _Init()
