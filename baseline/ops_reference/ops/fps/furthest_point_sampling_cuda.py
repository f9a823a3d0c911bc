"""`ops.fps.furthest_point_sampling_cuda` bound to the REFERENCE's own kernel.

The reference builds this module from ops/fps/src/{fps_api.cpp,sampling.cpp,sampling_gpu.cu}; sampling.cpp needs THC,
which torch no longer ships, so only that 15-line glue is replaced: the kernel + launcher of sampling_gpu.cu:24-184 are
compiled unmodified into oracle/_ref/libref_fps.so (oracle/Makefile) and called here with the tensors' pointers."""
import ctypes as C
import os

_REF = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "..", "..", "oracle", "_ref")
_lib = C.CDLL(os.path.join(os.path.abspath(_REF), "libref_fps.so"))
_lib.ref_fps_launch.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]


def furthest_point_sampling_wrapper(b, n, m, points_tensor, temp_tensor, idx_tensor):
    assert points_tensor.is_cuda and points_tensor.is_contiguous()       # CHECK_INPUT, sampling.cpp:9-21
    rc = _lib.ref_fps_launch(int(b), int(n), int(m), points_tensor.data_ptr(), temp_tensor.data_ptr(), idx_tensor.data_ptr())
    if rc != 0:
        raise RuntimeError("reference FPS kernel failed: CUDA error %d" % rc)
    return 1
