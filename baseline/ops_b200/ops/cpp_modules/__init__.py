"""`ops.cpp_modules` bound to librpcc_b200.so: the drop-in.  The reference's `from ops.cpp_modules import X` (dataset/
transformer.py:8, utils/segment_utils.py:8, utils/compress_utils.py:14-15, utils/contour_utils.py:5) resolves here."""
import sys

from rpcc_b200.plugin import (contour_utils_cpp, dataset_utils_cpp, feature_extractor_cpp, quantization_utils_cpp,  # noqa: F401
                              segment_utils_cpp)

for _m in (contour_utils_cpp, dataset_utils_cpp, feature_extractor_cpp, quantization_utils_cpp, segment_utils_cpp):
    sys.modules[__name__ + "." + _m.__name__.split(".")[-1]] = _m
