import ctypes as C

import numpy as np

from .. import _lib
from .._lib import check, ptr


def f32(a):
    # pybind's py::array_t<float> default is c_style | forcecast: wrong dtypes are converted by copy
    return np.ascontiguousarray(a, dtype=np.float32)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def hw(a):
    """(H, W) of a (H,W) or (H,W,1) array, as buffer_info.shape[0..1] in the reference."""
    return int(a.shape[0]), int(a.shape[1])


__all__ = ["C", "np", "_lib", "check", "ptr", "f32", "i32", "hw"]
