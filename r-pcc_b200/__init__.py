"""rpcc_b200 -- B200-native (sm_100a) implementation of R-PCC's per-frame compression hot path.

Host code is Python over a C-ABI CUDA library (include/rpcc_b200.h, r-pcc_b200/csrc).  The
package mirrors the reference's interface for this path: `plugin/*` are the drop-in
replacements of the reference's pybind modules; `dataset.py`, `segment_utils.py`, `compress_utils.py`,
`contour_utils.py`, `evaluate_metrics.py` mirror the L3 classes (PCTransformer, PointCloudSegment,
QuantizationModule, ...); `batch.py` holds the batched encoder / decoder, `hostio.py` the native host stage
(file reader, entropy-coder pool), `tools/` the four drivers.
There is no CPU fallback: every compute call needs librpcc_b200.so and a CUDA device.
"""
from ._lib import RpccError, build, launch_count, lib  # noqa: F401
from .lidar import LIDAR_TABLE, LidarConfig  # noqa: F401

__version__ = "0.1.0"
