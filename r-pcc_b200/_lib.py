"""ctypes loader for librpcc_b200.so (the C ABI declared in include/rpcc_b200.h).

There is no CPU fallback anywhere in this package: if the library is missing, or a compute
entry point is called without a CUDA device, the call raises."""
import ctypes as C
import os
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RPCC_B200_LIB") or os.path.join(_PKG, "lib", "librpcc_b200.so")  # env: A/B builds
_lib = None


class RpccError(RuntimeError):
    pass


def build(verbose=False):
    """Compile every CUDA source for sm_100a into lib/librpcc_b200.so (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", os.path.join(_PKG, "csrc"), "-j", str(min(16, os.cpu_count() or 4))]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RpccError("building librpcc_b200.so failed:\n" + res.stdout)
    if verbose:
        print(res.stdout)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RpccError("%s not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(there is no CPU fallback)" % LIB_PATH)
        _lib = C.CDLL(LIB_PATH)
        for name, rt in (("rpcc_last_error", C.c_char_p), ("rpcc_launch_count", C.c_longlong),
                         ("rpcc_encoder_stream", C.c_void_p), ("rpcc_encoder_device_buffer", C.c_void_p),
                         ("rpcc_book_bytes", C.c_size_t)):
            try:
                getattr(_lib, name).restype = rt
            except AttributeError as e:  # an incomplete build must not go unnoticed
                raise RpccError("librpcc_b200.so does not export %s (stale build?)" % name) from e
    return _lib


def check(rc):
    if rc != 0:
        raise RpccError("librpcc_b200 error %d: %s" % (rc, lib().rpcc_last_error().decode()))
    return rc


def ptr(a):
    """Address of a numpy array or torch tensor (or None)."""
    if a is None:
        return C.c_void_p(0)
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    return a.ctypes.data_as(C.c_void_p)


def launch_count():
    return int(lib().rpcc_launch_count())
