"""TEST INFRASTRUCTURE -- loaders for the reference's OWN compiled code in oracle/_ref.

oracle/_ref is built by oracle/Makefile from /root/reference (dev container only) and travels
to the GPU box prebuilt; it is never read by the product.  ``cpp(name)`` returns one of the
reference's five pybind11 modules (ops/cpp_modules/src/cpp_modules.cpp:597-636);
``fps()``/``chamfer()`` return ctypes handles to the reference CUDA kernels wrapped by
ref_fps_shim.cu / ref_chamfer_shim.cu (raw device pointers; they synchronise the device).
"""
import ctypes as C
import glob
import importlib.util
import os

_REF = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
_MODS = {}


def have_cpp():
    return bool(glob.glob(os.path.join(_REF, "ops", "cpp_modules", "dataset_utils_cpp*.so")))


def have_cuda():
    return os.path.exists(os.path.join(_REF, "libref_fps.so")) and \
        os.path.exists(os.path.join(_REF, "libref_chamfer.so"))


def cpp(name):
    if name not in _MODS:
        path = glob.glob(os.path.join(_REF, "ops", "cpp_modules", name + "*.so"))[0]
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _MODS[name] = mod
    return _MODS[name]


def fps():
    lib = C.CDLL(os.path.join(_REF, "libref_fps.so"))
    lib.ref_fps_launch.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    return lib


def chamfer():
    lib = C.CDLL(os.path.join(_REF, "libref_chamfer.so"))
    lib.ref_chamfer_launch.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    return lib
