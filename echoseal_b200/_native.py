"""ctypes loader for the C-ABI library (include/echoseal_b200.h).

There is NO CPU fallback: if `libechoseal_b200.so` is missing or CUDA is unavailable the product
path raises.  Build the library with `python -c "import __graft_entry__ as g; g.build()"` (nvcc,
sm_100a)."""
from __future__ import annotations
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# ES_B200_LIB: developer override used by tools/build_variants.py to A/B kernel variants; the product build is in-tree
LIB_PATH = os.environ.get("ES_B200_LIB") or os.path.join(_HERE, "libechoseal_b200.so")
_lib = None


class NativeError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeError(
                f"{LIB_PATH} not built — run __graft_entry__.build(); echoseal_b200 has no CPU fallback")
        L = C.CDLL(LIB_PATH)
        L.es_last_error.restype = C.c_char_p
        L.es_scl_scratch_bytes.restype = C.c_size_t
        _lib = L
    return _lib


LAUNCHES = 0        # kernels launched through the C-ABI since import (bench.py reports the delta)
KERNEL_TIMES = None  # when set to a dict, wrappers append (start_event, end_event) per kernel name


def count_launch(n: int = 1):
    global LAUNCHES
    LAUNCHES += n


class timed:
    """with timed("name"): <one launch>  — records CUDA events on the current (launching) stream when
    KERNEL_TIMES is enabled; otherwise only counts the launch."""
    def __init__(self, name, n=1):
        self.name, self.n = name, n
    def __enter__(self):
        count_launch(self.n)
        if KERNEL_TIMES is not None:
            import torch
            self.e0 = torch.cuda.Event(enable_timing=True); self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self
    def __exit__(self, *a):
        if KERNEL_TIMES is not None:
            self.e1.record()
            KERNEL_TIMES.setdefault(self.name, []).append((self.e0, self.e1))
        return False


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().es_last_error().decode("utf-8", "replace")
        raise NativeError(f"{what} failed rc={rc}: {msg}")


def ptr(t):
    """device (or host) pointer of a torch tensor / None."""
    if t is None:
        return C.c_void_p(0)
    return C.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(*tensors):
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise NativeError("tensor must live on a CUDA device (no CPU fallback)")
        if not t.is_contiguous():
            raise NativeError("tensor must be contiguous")
