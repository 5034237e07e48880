"""ctypes binding of libnmx.so (the C ABI declared in include/nmx.h).

There is NO fallback: if the shared library is missing or a call returns non-zero this raises.  Every op in
this package goes through `call(...)`, which passes raw device pointers + the current CUDA stream."""
import ctypes
import os
import re

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_lib", "libnmx.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "nmx.h")

_lib = None


class NmxError(RuntimeError):
    pass


def declared_symbols():
    """Names of all functions declared in include/nmx.h (used by the CPU-side export test)."""
    with open(HEADER_PATH) as f:
        txt = f.read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(nmx_[a-z0-9_]+)\s*\(", txt)))


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NmxError(
                f"{LIB_PATH} not found: the CUDA extension is not built. Run `python -m nerf_meets_mlx_b200.build` "
                "(or __graft_entry__.build()). There is no CPU fallback.")
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.nmx_last_error_string.restype = ctypes.c_char_p
        _lib.nmx_launch_count.restype = ctypes.c_int64
        _lib.nmx_version.restype = ctypes.c_int
        for name in ("nmx_mlp_param_count", "nmx_mlp_workspace_bytes"):
            if hasattr(_lib, name):
                getattr(_lib, name).restype = ctypes.c_int64
    return _lib


def ptr(t):
    if t is None:
        return ctypes.c_void_p(0)
    if isinstance(t, torch.Tensor):
        return ctypes.c_void_p(t.data_ptr())
    return ctypes.c_void_p(int(t))


def i64(v):
    return ctypes.c_int64(int(v))


def i32(v):
    return ctypes.c_int(int(v))


def f32(v):
    return ctypes.c_float(float(v))


def f64(v):
    return ctypes.c_double(float(v))


def stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def call(name, *args):
    fn = getattr(lib(), name)
    rc = fn(*args)
    if rc != 0:
        msg = lib().nmx_last_error_string().decode("utf-8", "replace")
        raise NmxError(f"{name} failed (code {rc}): {msg}")


def launch_count():
    return int(lib().nmx_launch_count())


def require_cuda(*tensors):
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise NmxError("nerf_meets_mlx_b200 ops need CUDA tensors (no CPU fallback exists)")
        if not t.is_contiguous():
            raise NmxError("nerf_meets_mlx_b200 ops need contiguous tensors")
