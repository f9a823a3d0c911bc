"""`ops.cpp_modules` bound to the REFERENCE's own C++: the five pybind modules compiled from the untouched
ops/cpp_modules/src/cpp_modules.cpp into oracle/_ref/ops/cpp_modules (oracle/Makefile)."""
import glob
import importlib.util
import os
import sys

_DIR = os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "..", "..", "oracle", "_ref",
                                    "ops", "cpp_modules"))
for _name in ("dataset_utils_cpp", "segment_utils_cpp", "quantization_utils_cpp", "feature_extractor_cpp", "contour_utils_cpp"):
    _path = glob.glob(os.path.join(_DIR, _name + "*.so"))[0]
    _spec = importlib.util.spec_from_file_location(_name, _path)
    _mod = importlib.util.module_from_spec(_spec)
    _spec.loader.exec_module(_mod)
    globals()[_name] = _mod
    sys.modules[__name__ + "." + _name] = _mod
