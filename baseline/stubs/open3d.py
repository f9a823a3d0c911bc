"""Stand-in for open3d (absent from this image, unpinned by the reference): only what the FPS path touches.

`PointCloud.segment_plane(distance_threshold, ransac_n, num_iterations)` is a vectorised numpy RANSAC following the
published algorithm (least-squares planes of `ransac_n` random points, inlier count at the threshold, ties
by rmse, refit on the inliers).  All hypotheses are fitted and scored in a handful of array operations (about 2 ms for
the 100 x 5000 ground problem of utils/segment_utils.py:101-108), so the stand-in is not slower than the multi-threaded
C++ it replaces by more than that.  The sampler is numpy's global RNG, unseeded like the reference's own subsample.
With RPCC_STUB_GROUND="a,b,c,d" in the environment the 10-point call (the ground fit, utils/segment_utils.py:79-81) returns
that model instead: the parity tests inject one ground plane into every implementation they compare."""
import os

import numpy as np


class _Vec(np.ndarray):
    pass


class utility:
    @staticmethod
    def Vector3dVector(a):
        return np.asarray(a, dtype=np.float64)


def _planes_from_sums(pts):
    """pts (I, n, 3) -> (I, 4) least-squares planes through each group (normal = smallest eigenvector of the scatter)."""
    c = pts.mean(1, keepdims=True)
    r = pts - c
    cov = np.einsum("inj,ink->ijk", r, r)
    w, v = np.linalg.eigh(cov)
    n = v[:, :, 0]
    d = -np.einsum("ij,ij->i", n, c[:, 0, :])
    return np.concatenate((n, d[:, None]), 1)


class _PointCloud:
    def __init__(self):
        self.points = np.zeros((0, 3))
        self.normals = None

    def segment_plane(self, distance_threshold=0.1, ransac_n=3, num_iterations=100):
        pts = np.asarray(self.points, np.float64).reshape(-1, 3)
        n = pts.shape[0]
        fixed = os.environ.get("RPCC_STUB_GROUND")
        if fixed and ransac_n == 10:
            return np.array([float(x) for x in fixed.split(",")]), []
        if n < ransac_n:
            return np.array([0.0, 0.0, 1.0, 0.0]), []
        idx = np.argsort(np.random.random((num_iterations, n)), axis=1)[:, :ransac_n] if n <= 64 else \
            np.random.randint(0, n, (num_iterations, ransac_n))
        planes = _planes_from_sums(pts[idx])
        dist = np.abs(pts @ planes[:, :3].T + planes[:, 3])            # (n, I)
        inl = dist < distance_threshold
        cnt = inl.sum(0)
        err = np.where(inl, dist * dist, 0.0).sum(0)
        rmse = np.sqrt(err / np.maximum(cnt, 1))
        best = np.lexsort((rmse, -cnt))[0]
        sel = np.where(inl[:, best])[0]
        model = planes[best]
        if sel.size >= 3:
            model = _planes_from_sums(pts[sel][None])[0]
        return model, sel.tolist()

    def estimate_normals(self, *a, **k):
        raise NotImplementedError("open3d stand-in: normals (point-to-plane PSNR) are outside this path")

    def cluster_dbscan(self, *a, **k):
        raise NotImplementedError("open3d stand-in: DBSCAN is outside this path")


class geometry:
    PointCloud = _PointCloud

    class KDTreeSearchParamHybrid:
        def __init__(self, *a, **k):
            pass


class io:
    @staticmethod
    def read_point_cloud(*a, **k):
        raise NotImplementedError("open3d stand-in: .ply/.pcd readers are outside this path")
