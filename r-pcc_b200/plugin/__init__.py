"""Drop-in replacements of the reference's native extension modules.

Same module names, function names, argument order/meaning and return shapes/dtypes as
`ops.cpp_modules.{dataset_utils_cpp, segment_utils_cpp, quantization_utils_cpp,
feature_extractor_cpp, contour_utils_cpp}` (ops/cpp_modules/src/cpp_modules.cpp:597-636),
`ops.fps.furthest_point_sampling_cuda` (ops/fps/src/fps_api.cpp:7-9) and the JIT `chamfer_3D`
(chamfer3D/chamfer_cuda.cpp:28-31); every function runs on the GPU through librpcc_b200.so.
INTEGRATION.md shows how the reference imports them."""
