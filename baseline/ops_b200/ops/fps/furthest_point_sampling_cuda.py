"""`ops.fps.furthest_point_sampling_cuda` bound to librpcc_b200.so (ops/fps/fps_utils.py:7 imports it relatively)."""
from rpcc_b200.plugin.furthest_point_sampling_cuda import furthest_point_sampling_wrapper  # noqa: F401
